/* Stand-in for the LV2 URID extension header (ABI-compatible subset). */
#ifndef STANDIN_LV2_URID_H
#define STANDIN_LV2_URID_H

#include <stdint.h>

#define LV2_URID_URI "http://lv2plug.in/ns/ext/urid"
#define LV2_URID_PREFIX LV2_URID_URI "#"
#define LV2_URID__map LV2_URID_PREFIX "map"
#define LV2_URID__unmap LV2_URID_PREFIX "unmap"

typedef void*    LV2_URID_Map_Handle;
typedef uint32_t LV2_URID;

typedef struct {
	LV2_URID_Map_Handle handle;
	LV2_URID (*map) (LV2_URID_Map_Handle handle, const char* uri);
} LV2_URID_Map;

#endif
