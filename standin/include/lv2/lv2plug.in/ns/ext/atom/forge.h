/* Stand-in for the LV2 atom forge header: buffer-backed forge only (no sink
 * callbacks), enough for plugins that write an atom sequence into a host
 * provided port buffer.  Pulls in <assert.h>/<stdbool.h> like the original,
 * which src/phaserotate.c of the reference relies on. */
#ifndef STANDIN_LV2_ATOM_FORGE_H
#define STANDIN_LV2_ATOM_FORGE_H

#include <assert.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "../urid/urid.h"
#include "atom.h"
#include "util.h"

typedef void*    LV2_Atom_Forge_Sink_Handle;
typedef intptr_t LV2_Atom_Forge_Ref;
typedef LV2_Atom_Forge_Ref (*LV2_Atom_Forge_Sink) (LV2_Atom_Forge_Sink_Handle handle, const void* buf, uint32_t size);
typedef LV2_Atom* (*LV2_Atom_Forge_Deref_Func) (LV2_Atom_Forge_Sink_Handle handle, LV2_Atom_Forge_Ref ref);

typedef struct LV2_Atom_Forge_Frame {
	struct LV2_Atom_Forge_Frame* parent;
	LV2_Atom_Forge_Ref           ref;
} LV2_Atom_Forge_Frame;

typedef struct {
	uint8_t* buf;
	uint32_t offset;
	uint32_t size;

	LV2_Atom_Forge_Sink        sink;
	LV2_Atom_Forge_Deref_Func  deref;
	LV2_Atom_Forge_Sink_Handle handle;

	LV2_Atom_Forge_Frame* stack;

	LV2_URID Blank;
	LV2_URID Bool;
	LV2_URID Chunk;
	LV2_URID Double;
	LV2_URID Float;
	LV2_URID Int;
	LV2_URID Long;
	LV2_URID Literal;
	LV2_URID Object;
	LV2_URID Path;
	LV2_URID Property;
	LV2_URID Resource;
	LV2_URID Sequence;
	LV2_URID String;
	LV2_URID Tuple;
	LV2_URID URI;
	LV2_URID URID;
	LV2_URID Vector;
} LV2_Atom_Forge;

static inline void
lv2_atom_forge_set_buffer (LV2_Atom_Forge* forge, uint8_t* buf, size_t size)
{
	forge->buf    = buf;
	forge->size   = (uint32_t)size;
	forge->offset = 0;
	forge->deref  = NULL;
	forge->sink   = NULL;
	forge->handle = NULL;
	forge->stack  = NULL;
}

static inline void
lv2_atom_forge_init (LV2_Atom_Forge* forge, LV2_URID_Map* map)
{
	lv2_atom_forge_set_buffer (forge, NULL, 0);
	forge->Blank    = map->map (map->handle, LV2_ATOM__Blank);
	forge->Bool     = map->map (map->handle, LV2_ATOM__Bool);
	forge->Chunk    = map->map (map->handle, LV2_ATOM__Chunk);
	forge->Double   = map->map (map->handle, LV2_ATOM__Double);
	forge->Float    = map->map (map->handle, LV2_ATOM__Float);
	forge->Int      = map->map (map->handle, LV2_ATOM__Int);
	forge->Long     = map->map (map->handle, LV2_ATOM__Long);
	forge->Literal  = map->map (map->handle, LV2_ATOM__Literal);
	forge->Object   = map->map (map->handle, LV2_ATOM__Object);
	forge->Path     = map->map (map->handle, LV2_ATOM__Path);
	forge->Property = map->map (map->handle, LV2_ATOM__Property);
	forge->Resource = map->map (map->handle, LV2_ATOM__Resource);
	forge->Sequence = map->map (map->handle, LV2_ATOM__Sequence);
	forge->String   = map->map (map->handle, LV2_ATOM__String);
	forge->Tuple    = map->map (map->handle, LV2_ATOM__Tuple);
	forge->URI      = map->map (map->handle, LV2_ATOM__URI);
	forge->URID     = map->map (map->handle, LV2_ATOM__URID);
	forge->Vector   = map->map (map->handle, LV2_ATOM__Vector);
}

static inline LV2_Atom*
lv2_atom_forge_deref (LV2_Atom_Forge* forge, LV2_Atom_Forge_Ref ref)
{
	(void)forge;
	return (LV2_Atom*)ref;
}

/* append raw bytes; every open container grows by the same amount */
static inline LV2_Atom_Forge_Ref
lv2_atom_forge_raw (LV2_Atom_Forge* forge, const void* data, uint32_t size)
{
	if (!forge->buf || forge->offset + size > forge->size) {
		return 0;
	}
	uint8_t* mem = forge->buf + forge->offset;
	forge->offset += size;
	memcpy (mem, data, size);
	for (LV2_Atom_Forge_Frame* f = forge->stack; f; f = f->parent) {
		lv2_atom_forge_deref (forge, f->ref)->size += size;
	}
	return (LV2_Atom_Forge_Ref)mem;
}

static inline void
lv2_atom_forge_pad (LV2_Atom_Forge* forge, uint32_t written)
{
	const uint64_t pad      = 0;
	const uint32_t pad_size = lv2_atom_pad_size (written) - written;
	lv2_atom_forge_raw (forge, &pad, pad_size);
}

static inline LV2_Atom_Forge_Ref
lv2_atom_forge_write (LV2_Atom_Forge* forge, const void* data, uint32_t size)
{
	LV2_Atom_Forge_Ref out = lv2_atom_forge_raw (forge, data, size);
	if (out) {
		lv2_atom_forge_pad (forge, size);
	}
	return out;
}

static inline LV2_Atom_Forge_Ref
lv2_atom_forge_push (LV2_Atom_Forge* forge, LV2_Atom_Forge_Frame* frame, LV2_Atom_Forge_Ref ref)
{
	frame->parent = forge->stack;
	frame->ref    = ref;
	if (ref) {
		forge->stack = frame;
	}
	return ref;
}

static inline void
lv2_atom_forge_pop (LV2_Atom_Forge* forge, LV2_Atom_Forge_Frame* frame)
{
	if (frame->ref) {
		assert (frame == forge->stack);
		forge->stack = frame->parent;
	}
}

static inline LV2_Atom_Forge_Ref
lv2_atom_forge_primitive (LV2_Atom_Forge* forge, const LV2_Atom* a)
{
	return lv2_atom_forge_write (forge, a, (uint32_t)sizeof (LV2_Atom) + a->size);
}

static inline LV2_Atom_Forge_Ref
lv2_atom_forge_int (LV2_Atom_Forge* forge, int32_t val)
{
	const LV2_Atom_Int a = { { sizeof (val), forge->Int }, val };
	return lv2_atom_forge_primitive (forge, &a.atom);
}

static inline LV2_Atom_Forge_Ref
lv2_atom_forge_float (LV2_Atom_Forge* forge, float val)
{
	const LV2_Atom_Float a = { { sizeof (val), forge->Float }, val };
	return lv2_atom_forge_primitive (forge, &a.atom);
}

static inline LV2_Atom_Forge_Ref
lv2_atom_forge_bool (LV2_Atom_Forge* forge, bool val)
{
	const LV2_Atom_Bool a = { { sizeof (int32_t), forge->Bool }, val ? 1 : 0 };
	return lv2_atom_forge_primitive (forge, &a.atom);
}

static inline LV2_Atom_Forge_Ref
lv2_atom_forge_object (LV2_Atom_Forge* forge, LV2_Atom_Forge_Frame* frame, LV2_URID id, LV2_URID otype)
{
	const LV2_Atom_Object a = { { (uint32_t)sizeof (LV2_Atom_Object_Body), forge->Object }, { id, otype } };
	return lv2_atom_forge_push (forge, frame, lv2_atom_forge_write (forge, &a, (uint32_t)sizeof (a)));
}

/* pre-1.8 spelling: same layout, atom type Blank */
static inline LV2_Atom_Forge_Ref
lv2_atom_forge_blank (LV2_Atom_Forge* forge, LV2_Atom_Forge_Frame* frame, uint32_t id, LV2_URID otype)
{
	const LV2_Atom_Object a = { { (uint32_t)sizeof (LV2_Atom_Object_Body), forge->Blank }, { id, otype } };
	return lv2_atom_forge_push (forge, frame, lv2_atom_forge_write (forge, &a, (uint32_t)sizeof (a)));
}

static inline LV2_Atom_Forge_Ref
lv2_atom_forge_property_head (LV2_Atom_Forge* forge, LV2_URID key, LV2_URID context)
{
	const uint32_t head[2] = { key, context };
	return lv2_atom_forge_raw (forge, head, (uint32_t)sizeof (head));
}

static inline LV2_Atom_Forge_Ref
lv2_atom_forge_sequence_head (LV2_Atom_Forge* forge, LV2_Atom_Forge_Frame* frame, uint32_t unit)
{
	const LV2_Atom_Sequence a = { { (uint32_t)sizeof (LV2_Atom_Sequence_Body), forge->Sequence }, { unit, 0 } };
	return lv2_atom_forge_push (forge, frame, lv2_atom_forge_write (forge, &a, (uint32_t)sizeof (a)));
}

static inline LV2_Atom_Forge_Ref
lv2_atom_forge_frame_time (LV2_Atom_Forge* forge, int64_t frames)
{
	return lv2_atom_forge_write (forge, &frames, (uint32_t)sizeof (frames));
}

#endif
