/*
 * Stand-in for the LV2 core header (LV2 is not installed in this image).
 * ABI-compatible subset of the LV2 core C API: descriptor, feature, handle.
 * Used to compile src/phaserotate.c of the reference unmodified and this
 * repository's own plugin shim.  A deployment uses the real LV2 headers.
 */
#ifndef STANDIN_LV2_H
#define STANDIN_LV2_H

#include <stdint.h>

#define LV2_CORE_URI "http://lv2plug.in/ns/lv2core"
#define LV2_CORE_PREFIX LV2_CORE_URI "#"

#ifdef __cplusplus
extern "C" {
#endif

typedef void* LV2_Handle;

typedef struct {
	const char* URI;
	void*       data;
} LV2_Feature;

typedef struct LV2_Descriptor {
	const char* URI;
	LV2_Handle (*instantiate) (const struct LV2_Descriptor* descriptor,
	                           double                       sample_rate,
	                           const char*                  bundle_path,
	                           const LV2_Feature* const*    features);
	void (*connect_port) (LV2_Handle instance, uint32_t port, void* data_location);
	void (*activate) (LV2_Handle instance);
	void (*run) (LV2_Handle instance, uint32_t sample_count);
	void (*deactivate) (LV2_Handle instance);
	void (*cleanup) (LV2_Handle instance);
	const void* (*extension_data) (const char* uri);
} LV2_Descriptor;

#ifdef _WIN32
#define LV2_SYMBOL_EXPORT __declspec(dllexport)
#else
#define LV2_SYMBOL_EXPORT __attribute__ ((visibility ("default")))
#endif

LV2_SYMBOL_EXPORT
const LV2_Descriptor* lv2_descriptor (uint32_t index);

typedef const LV2_Descriptor* (*LV2_Descriptor_Function) (uint32_t index);

#ifdef __cplusplus
}
#endif
#endif
