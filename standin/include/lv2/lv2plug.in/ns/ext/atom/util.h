/* Stand-in for the LV2 atom utility header: sequence iteration and object
 * property lookup, written against the LV2 atom memory layout. */
#ifndef STANDIN_LV2_ATOM_UTIL_H
#define STANDIN_LV2_ATOM_UTIL_H

#include <stdarg.h>
#include <stdbool.h>
#include <stdint.h>
#include <string.h>

#include "atom.h"

static inline uint32_t
lv2_atom_pad_size (uint32_t size)
{
	return (size + 7U) & ~7U;
}

static inline uint32_t
lv2_atom_total_size (const LV2_Atom* atom)
{
	return (uint32_t)sizeof (LV2_Atom) + atom->size;
}

static inline LV2_Atom_Event*
lv2_atom_sequence_begin (const LV2_Atom_Sequence_Body* body)
{
	return (LV2_Atom_Event*)(body + 1);
}

static inline bool
lv2_atom_sequence_is_end (const LV2_Atom_Sequence_Body* body, uint32_t size, const LV2_Atom_Event* i)
{
	return (const uint8_t*)i >= ((const uint8_t*)body + size);
}

static inline LV2_Atom_Event*
lv2_atom_sequence_next (const LV2_Atom_Event* i)
{
	return (LV2_Atom_Event*)((const uint8_t*)i + sizeof (LV2_Atom_Event) + lv2_atom_pad_size (i->body.size));
}

/* lv2_atom_object_get (obj, key0, &atom0, key1, &atom1, ..., 0) */
static inline int
lv2_atom_object_get (const LV2_Atom_Object* object, ...)
{
	int     matches = 0;
	const uint8_t* const body_end = (const uint8_t*)&object->body + object->atom.size;
	const uint8_t* p = (const uint8_t*)(&object->body + 1);
	while (p + sizeof (LV2_Atom_Property_Body) <= body_end) {
		const LV2_Atom_Property_Body* prop = (const LV2_Atom_Property_Body*)p;
		va_list args;
		va_start (args, object);
		for (;;) {
			const uint32_t key = va_arg (args, uint32_t);
			if (!key) {
				break;
			}
			const LV2_Atom** dst = va_arg (args, const LV2_Atom**);
			if (key == prop->key && !*dst) {
				*dst = &prop->value;
				++matches;
				break;
			}
		}
		va_end (args);
		p += lv2_atom_pad_size ((uint32_t)sizeof (LV2_Atom_Property_Body) + prop->value.size);
	}
	return matches;
}

#endif
