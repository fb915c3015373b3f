/* Stand-in for the LV2 options extension header (ABI-compatible subset). */
#ifndef STANDIN_LV2_OPTIONS_H
#define STANDIN_LV2_OPTIONS_H

#include <stdint.h>
#include "../urid/urid.h"

#define LV2_OPTIONS_URI "http://lv2plug.in/ns/ext/options"
#define LV2_OPTIONS_PREFIX LV2_OPTIONS_URI "#"
#define LV2_OPTIONS__options LV2_OPTIONS_PREFIX "options"

typedef enum {
	LV2_OPTIONS_INSTANCE,
	LV2_OPTIONS_RESOURCE,
	LV2_OPTIONS_BLANK,
	LV2_OPTIONS_PORT
} LV2_Options_Context;

typedef struct {
	LV2_Options_Context context;
	uint32_t            subject;
	LV2_URID            key;
	uint32_t            size;
	LV2_URID            type;
	const void*         value;
} LV2_Options_Option;

#endif
