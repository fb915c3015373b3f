/* Stand-in for the LV2 log extension convenience header.  The reference
 * includes it (src/phaserotate.c:37) but uses nothing from it. */
#ifndef STANDIN_LV2_LOGGER_H
#define STANDIN_LV2_LOGGER_H
#define LV2_LOG_URI "http://lv2plug.in/ns/ext/log"
#endif
