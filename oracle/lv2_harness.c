/*
 * TEST INFRASTRUCTURE — a minimal LV2 host.
 *
 * dlopen()s an LV2 plugin binary that exports lv2_descriptor() with the
 * phaserotate port layout (reference: src/phaserotate.h:99-111; ports
 * 0 control, 1 notify, 2 latency, then per channel angle/in/out) and drives
 * run() the way a host does: urid:map feature, an empty control sequence and a
 * sized notify buffer on ports 0/1 (src/phaserotate.c:790 returns early
 * without them), fixed-size calls, a per-call angle schedule.
 *
 * The same harness drives the reference plugin built from the unmodified
 * source (oracle/_ref/phaserotate_ref.so) and this repository's CUDA-backed
 * plugin, so parity tests feed both identical calls.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <lv2/lv2plug.in/ns/ext/atom/atom.h>
#include <lv2/lv2plug.in/ns/ext/urid/urid.h>
#include <lv2/lv2plug.in/ns/lv2core/lv2.h>

#define MAX_URIS 256
static char*    g_uris[MAX_URIS];
static uint32_t g_n_uris = 0;

static LV2_URID
map_uri (LV2_URID_Map_Handle h, const char* uri)
{
	(void)h;
	for (uint32_t i = 0; i < g_n_uris; ++i) {
		if (!strcmp (g_uris[i], uri)) {
			return i + 1;
		}
	}
	if (g_n_uris >= MAX_URIS) {
		return 0;
	}
	g_uris[g_n_uris] = strdup (uri);
	return ++g_n_uris;
}

/*
 * in / out: planar [n_chn][n_frames]; angles: [n_calls][n_chn] degrees, the
 * value on the angle port for each run() call (n_calls = ceil(n_frames/block)).
 * inplace != 0 connects the output buffer as input too.
 * Returns seconds spent inside run() calls, < 0 on error; *latency_out receives
 * the value the plugin wrote to its latency port.
 */
double
lv2h_render (const char* so_path, int index, double rate, int n_chn,
             const float* in, float* out, int64_t n_frames, uint32_t block,
             const float* angles, int inplace, float* latency_out)
{
	void* dl = dlopen (so_path, RTLD_NOW | RTLD_LOCAL);
	if (!dl) {
		fprintf (stderr, "lv2h: dlopen: %s\n", dlerror ());
		return -1;
	}
	LV2_Descriptor_Function df = (LV2_Descriptor_Function)dlsym (dl, "lv2_descriptor");
	if (!df) {
		dlclose (dl);
		return -2;
	}
	const LV2_Descriptor* d = df ((uint32_t)index);
	if (!d) {
		dlclose (dl);
		return -3;
	}

	LV2_URID_Map       map      = { NULL, map_uri };
	LV2_Feature        map_feat = { LV2_URID__map, &map };
	const LV2_Feature* feats[]  = { &map_feat, NULL };

	LV2_Handle h = d->instantiate (d, rate, "/tmp/", feats);
	if (!h) {
		dlclose (dl);
		return -4;
	}

	/* atom ports: 8-byte aligned, notify sized per the ttl minimum (4096) x2 */
	uint64_t           control_mem[4];
	uint64_t           notify_mem[1024 + 4];
	LV2_Atom_Sequence* control = (LV2_Atom_Sequence*)control_mem;
	LV2_Atom_Sequence* notify  = (LV2_Atom_Sequence*)notify_mem;
	const LV2_URID     seq     = map_uri (NULL, LV2_ATOM__Sequence);
	const LV2_URID     chunk   = map_uri (NULL, LV2_ATOM__Chunk);

	float  latency = -1;
	float  angle[2] = { 0, 0 };
	float* ibuf[2];
	float* obuf[2];
	for (int c = 0; c < n_chn; ++c) {
		ibuf[c] = (float*)malloc (sizeof (float) * block);
		obuf[c] = inplace ? ibuf[c] : (float*)malloc (sizeof (float) * block);
	}

	d->connect_port (h, 0, control);
	d->connect_port (h, 1, notify);
	d->connect_port (h, 2, &latency);
	for (int c = 0; c < n_chn; ++c) {
		d->connect_port (h, 3 + 3 * c, &angle[c]);
		d->connect_port (h, 4 + 3 * c, ibuf[c]);
		d->connect_port (h, 5 + 3 * c, obuf[c]);
	}
	if (d->activate) {
		d->activate (h);
	}

	double  spent = 0;
	int64_t call  = 0;
	for (int64_t pos = 0; pos < n_frames; pos += block, ++call) {
		const uint32_t n = (uint32_t)((n_frames - pos) < (int64_t)block ? (n_frames - pos) : (int64_t)block);
		control->atom.size = sizeof (LV2_Atom_Sequence_Body);
		control->atom.type = seq;
		control->body.unit = 0;
		control->body.pad  = 0;
		notify->atom.size  = 8192;
		notify->atom.type  = chunk;
		for (int c = 0; c < n_chn; ++c) {
			angle[c] = angles[call * n_chn + c];
			memcpy (ibuf[c], in + (size_t)c * n_frames + pos, sizeof (float) * n);
		}
		struct timespec t0, t1;
		clock_gettime (CLOCK_MONOTONIC, &t0);
		d->run (h, n);
		clock_gettime (CLOCK_MONOTONIC, &t1);
		spent += (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
		for (int c = 0; c < n_chn; ++c) {
			memcpy (out + (size_t)c * n_frames + pos, obuf[c], sizeof (float) * n);
		}
	}

	if (d->deactivate) {
		d->deactivate (h);
	}
	d->cleanup (h);
	for (int c = 0; c < n_chn; ++c) {
		if (!inplace) {
			free (obuf[c]);
		}
		free (ibuf[c]);
	}
	if (latency_out) {
		*latency_out = latency;
	}
	dlclose (dl);
	return spent;
}
