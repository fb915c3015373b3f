// LV2 plugin "http://gareus.org/oss/lv2/phaserotate" (mono, index 0) and
// "...#stereo" (index 1) with the DSP running in libphaserot_cuda.
//
// The LV2 boundary is the reference's, unchanged: same URIs, same port indices
// (src/phaserotate.h:99-111), same descriptor callbacks (src/phaserotate.c:860-893),
// same atom protocol with the GUI (ui_on / ui_off / reset_peaks / state in,
// state / levels out; src/phaserotate.c:523-536, 741-771, 801-830), same
// reported latency.  lv2ttl/*.in of the reference describe this binary as is.
//
// What moved to the GPU is the audio path of process_channel()
// (src/phaserotate.c:615-721): one phaserot_process_levels() call per run(), which
// also returns the per-call maxima of the delayed input and of the output,
// reduced on the device; only the meter ballistics (per call) stay on the host.
// No CPU DSP fallback: if the CUDA backend cannot be created, instantiate()
// returns NULL like the reference does on any allocation failure (src:315-372).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#ifdef HAVE_LV2_1_18_6
#include <lv2/atom/atom.h>
#include <lv2/atom/forge.h>
#include <lv2/core/lv2.h>
#include <lv2/options/options.h>
#include <lv2/urid/urid.h>
#else
#include <lv2/lv2plug.in/ns/ext/atom/atom.h>
#include <lv2/lv2plug.in/ns/ext/atom/forge.h>
#include <lv2/lv2plug.in/ns/ext/options/options.h>
#include <lv2/lv2plug.in/ns/ext/urid/urid.h>
#include <lv2/lv2plug.in/ns/lv2core/lv2.h>
#endif

#include "phaserot_cuda.h"

#define PLUGIN_URI "http://gareus.org/oss/lv2/phaserotate"
#define NS PLUGIN_URI "#"

namespace {

constexpr uint32_t kMaxChannels = 2; // src/phaserotate.h:97

enum Port : uint32_t { // src/phaserotate.h:99-111
	kControl = 0,
	kNotify  = 1,
	kLatency = 2,
	kAngle0  = 3,
	kInput0  = 4,
	kOutput0 = 5,
};

struct Uris {
	LV2_URID blank, object, vfloat, vint;
	LV2_URID ui_on, ui_off, reset_peaks, state, s_uiscale, s_link;
	LV2_URID levels, l_channel, l_in_cur, l_in_mom, l_in_peak, l_out_cur, l_out_mom, l_out_peak, l_diff_cur, l_diff_min, l_diff_max;

	void map_all (LV2_URID_Map* m)
	{
		auto u      = [m] (const char* s) { return m->map (m->handle, s); };
		blank       = u (LV2_ATOM__Blank);
		object      = u (LV2_ATOM__Object);
		(void)u (LV2_ATOM__Vector); // mapped by the reference too; keeps URID numbering identical for hosts that care
		vfloat      = u (LV2_ATOM__Float);
		vint        = u (LV2_ATOM__Int);
		(void)u (LV2_ATOM__eventTransfer);
		ui_on       = u (NS "ui_on");
		ui_off      = u (NS "ui_off");
		reset_peaks = u (NS "reset_peaks");
		state       = u (NS "state");
		s_uiscale   = u (NS "uiscale");
		s_link      = u (NS "link");
		levels      = u (NS "levels");
		l_channel   = u (NS "l_channel");
		l_in_cur    = u (NS "l_in_cur");
		l_in_mom    = u (NS "l_in_mom");
		l_in_peak   = u (NS "l_in_peak");
		l_out_cur   = u (NS "l_out_cur");
		l_out_mom   = u (NS "l_out_mom");
		l_out_peak  = u (NS "l_out_peak");
		l_diff_cur  = u (NS "l_diff_cur");
		l_diff_min  = u (NS "l_diff_min");
		l_diff_max  = u (NS "l_diff_max");
	}
};

// One level meter: running peak, momentary value with hold and 15 dB/s release
// (behaviour of meter_proc, src/phaserotate.c:451-470).
struct Meter {
	float peak = 0, momentary = 0;
	int   hold = 0;

	float feed (float level, uint32_t hold_frames, uint32_t frames_per_period, float falloff)
	{
		if (!std::isfinite (level)) {
			level = 0;
		}
		peak = std::fmax (peak, level);
		if (level > momentary) {
			momentary = level;
			hold      = (int)hold_frames;
		} else if (hold > 0) {
			hold -= (int)frames_per_period;
		} else {
			momentary = momentary * falloff + 1e-20f;
		}
		return level;
	}
	void clear ()
	{
		peak = momentary = 0;
	}
};

struct ChannelState {
	float* in    = nullptr;
	float* out   = nullptr;
	float* angle = nullptr;

	Meter              m_in, m_out;
	float              diff_min = 1, diff_max = 1;
	int                reset_delay = 0;
	float              last_target = 0; // for "angle changed" detection (meter reset, src:611)

	void reset_meters ()
	{
		m_in.clear ();
		m_out.clear ();
		diff_min = diff_max = 1;
	}
};

struct Plugin {
	const LV2_Atom_Sequence* control = nullptr;
	LV2_Atom_Sequence*       notify  = nullptr;
	float*                   latency_port = nullptr;

	uint32_t n_chn   = 1;
	float    rate    = 48000;
	uint32_t latency = 0;

	uint32_t hold_frames = 0, period = 0;
	float    falloff     = 0;

	LV2_URID_Map*        map = nullptr;
	Uris                 uris {};
	LV2_Atom_Forge       forge {};
	LV2_Atom_Forge_Frame seq_frame {};

	bool  ui_active = false, send_state = false, link = false;
	float ui_scale  = 1.f;

	phaserot_t*  dsp = nullptr;
	ChannelState ch[kMaxChannels];
	float        target_state[kMaxChannels] = { 0, 0 }; // angle state mirror, in turns
};

LV2_Atom_Forge_Ref
open_object (Plugin* p, LV2_Atom_Forge_Frame* f, LV2_URID otype)
{
#ifdef HAVE_LV2_1_8
	return lv2_atom_forge_object (&p->forge, f, 1, otype);
#else
	return lv2_atom_forge_blank (&p->forge, f, 1, otype); // what the reference emits without HAVE_LV2_1_8 (src/phaserotate.h:35-39)
#endif
}

void
put_float (Plugin* p, LV2_URID key, float v)
{
	lv2_atom_forge_property_head (&p->forge, key, 0);
	lv2_atom_forge_float (&p->forge, v);
}

void
send_state (Plugin* p) // src:522-536
{
	LV2_Atom_Forge_Frame f;
	lv2_atom_forge_frame_time (&p->forge, 0);
	open_object (p, &f, p->uris.state);
	put_float (p, p->uris.s_uiscale, p->ui_scale);
	lv2_atom_forge_property_head (&p->forge, p->uris.s_link, 0);
	lv2_atom_forge_bool (&p->forge, p->link);
	lv2_atom_forge_pop (&p->forge, &f);
}

void
cleanup (LV2_Handle instance)
{
	Plugin* p = (Plugin*)instance;
	if (!p) {
		return;
	}
	phaserot_destroy (p->dsp);
	delete p;
}

LV2_Handle
instantiate (const LV2_Descriptor* descriptor, double rate, const char*, const LV2_Feature* const* features)
{
	Plugin* p = new (std::nothrow) Plugin ();
	if (!p) {
		return nullptr;
	}
	if (!strcmp (descriptor->URI, PLUGIN_URI)) {
		p->n_chn = 1;
	} else if (!strcmp (descriptor->URI, PLUGIN_URI "#stereo")) {
		p->n_chn = 2;
	} else {
		delete p;
		return nullptr;
	}
	const LV2_Options_Option* options = nullptr;
	for (int i = 0; features && features[i]; ++i) {
		if (!strcmp (features[i]->URI, LV2_URID__map)) {
			p->map = (LV2_URID_Map*)features[i]->data;
		} else if (!strcmp (features[i]->URI, LV2_OPTIONS__options)) {
			options = (const LV2_Options_Option*)features[i]->data;
		}
	}
	if (!p->map) {
		fprintf (stderr, "phaserotate.lv2 error: Host does not support urid:map\n");
		delete p;
		return nullptr;
	}
	lv2_atom_forge_init (&p->forge, p->map);
	p->uris.map_all (p->map);

	float opt_scale = 0.f;
	if (options) {
		const LV2_URID atom_float = p->map->map (p->map->handle, LV2_ATOM__Float);
		const LV2_URID scale_key  = p->map->map (p->map->handle, "http://lv2plug.in/ns/extensions/ui#scaleFactor");
		for (const LV2_Options_Option* o = options; o->key; ++o) {
			if (o->context == LV2_OPTIONS_INSTANCE && o->key == scale_key && o->type == atom_float) {
				opt_scale = std::fmin (2.f, std::fmax (1.f, *(const float*)o->value));
			}
		}
	}
	(void)opt_scale;     // the reference overwrites the option with 1.0 a few lines later (src:273 vs 300)
	p->ui_scale = 1.0f;

	p->rate        = (float)rate;
	p->hold_frames = (uint32_t)(0.5 * rate + 0.5f); // src:303

	phaserot_cfg_t cfg;
	memset (&cfg, 0, sizeof (cfg));
	cfg.abi_version = PHASEROT_ABI_VERSION;
	cfg.mode        = PHASEROT_MODE_PLUGIN;
	cfg.n_channels  = (int32_t)p->n_chn;
	cfg.sample_rate = rate;
	cfg.device      = -1;
	const int rc    = phaserot_create (&p->dsp, &cfg);
	if (rc != PHASEROT_OK) {
		fprintf (stderr, "phaserotate.lv2 error: CUDA backend unavailable: %s (%s)\n", phaserot_strerror (rc), phaserot_last_error ());
		delete p;
		return nullptr;
	}
	p->latency = phaserot_latency (p->dsp);
	return (LV2_Handle)p;
}

void
connect_port (LV2_Handle instance, uint32_t port, void* data)
{
	Plugin* p = (Plugin*)instance;
	switch (port) {
		case kControl: p->control = (const LV2_Atom_Sequence*)data; return;
		case kNotify: p->notify = (LV2_Atom_Sequence*)data; return;
		case kLatency: p->latency_port = (float*)data; return;
		default: break;
	}
	const uint32_t c = (port - kAngle0) / 3;
	if (port < kAngle0 || c >= kMaxChannels) {
		return;
	}
	switch (kAngle0 + (port - kAngle0) % 3) {
		case kAngle0: p->ch[c].angle = (float*)data; break;
		case kInput0: p->ch[c].in = (float*)data; break;
		case kOutput0: p->ch[c].out = (float*)data; break;
	}
}

void
activate (LV2_Handle instance)
{
	Plugin* p = (Plugin*)instance;
	phaserot_reset (p->dsp);
	for (uint32_t c = 0; c < p->n_chn; ++c) {
		p->ch[c].reset_meters ();
		p->ch[c].reset_delay = (int)p->latency;
	}
}

void
run (LV2_Handle instance, uint32_t n_samples)
{
	Plugin* p = (Plugin*)instance;

	// forward the dry signal when not running in place (src:780-785): it is what
	// the host hears if run() returns early below
	for (uint32_t c = 0; c < p->n_chn; ++c) {
		if (p->ch[c].in != p->ch[c].out) {
			memcpy (p->ch[c].out, p->ch[c].in, sizeof (float) * n_samples);
		}
	}
	*p->latency_port = (float)p->latency;
	if (!p->control || !p->notify) {
		return; // latency measurement callback (src:790-793)
	}

	const uint32_t capacity = p->notify->atom.size;
	lv2_atom_forge_set_buffer (&p->forge, (uint8_t*)p->notify, capacity);
	lv2_atom_forge_sequence_head (&p->forge, &p->seq_frame, 0);

	// messages from the GUI (src:801-830)
	for (LV2_Atom_Event* ev = lv2_atom_sequence_begin (&p->control->body);
	     !lv2_atom_sequence_is_end (&p->control->body, p->control->atom.size, ev);
	     ev = lv2_atom_sequence_next (ev)) {
		if (ev->body.type != p->uris.blank && ev->body.type != p->uris.object) {
			continue;
		}
		const LV2_Atom_Object* obj = (const LV2_Atom_Object*)&ev->body;
		const LV2_URID         t   = obj->body.otype;
		if (t == p->uris.ui_off) {
			p->ui_active = false;
		} else if (t == p->uris.ui_on) {
			p->ui_active  = true;
			p->send_state = true;
		} else if (t == p->uris.reset_peaks) {
			for (uint32_t c = 0; c < p->n_chn; ++c) {
				p->ch[c].reset_meters ();
			}
		} else if (t == p->uris.state) {
			const LV2_Atom* v = nullptr;
			lv2_atom_object_get (obj, p->uris.s_uiscale, &v, 0);
			if (v) {
				p->ui_scale = ((const LV2_Atom_Float*)v)->body;
			}
			v = nullptr;
			lv2_atom_object_get (obj, p->uris.s_link, &v, 0);
			if (v) {
				p->link = ((const LV2_Atom_Bool*)v)->body != 0;
			}
		}
	}

	if (p->period != n_samples) { // 15 dB/s release, per period (src:832-838)
		p->falloff = powf (10.0f, -0.05f * 15.0f * ((float)n_samples / p->rate));
		p->period  = n_samples;
	}

	// --- audio: all channels in one backend call --------------------------
	float        lvl_in[kMaxChannels] = { 0, 0 };
	float        angles[kMaxChannels] = { 0, 0 };
	// the backend's ramp state as process_channel() finds it at the start of this call (Channel::angle)
	float        angle_state[kMaxChannels] = { 0, 0 };
	if (phaserot_plugin_angle (p->dsp, angle_state) != PHASEROT_OK) {
		angle_state[0] = p->target_state[0];
		angle_state[1] = p->target_state[1];
	}
	const float* ins[kMaxChannels]    = { nullptr, nullptr };
	float*       outs[kMaxChannels]   = { nullptr, nullptr };
	for (uint32_t c = 0; c < p->n_chn; ++c) {
		ChannelState& ch = p->ch[c];
		angles[c]        = *ch.angle;
		ins[c]           = ch.out; // processed in place on the output buffer like the reference (src:563)
		outs[c]          = ch.out;
		// meter_delayed_reset (src:497-509, 611): after an angle change the
		// output/diff meters restart once the new setting has reached the output
		float target = std::fmin (0.5f, std::fmax (-0.5f, angles[c] / -360.f));
		if (ch.reset_delay > 0) {
			ch.diff_min = ch.diff_max = 1;
			ch.m_out.momentary        = 0;
			ch.reset_delay -= (int)n_samples;
		}
		if (target != angle_state[c]) { // src:611: re-armed on every call while the ramp has not arrived
			ch.reset_delay = (int)(p->latency + n_samples);
		}
		ch.last_target = target;
	}
	// the input level is measured `latency` samples late so that it lines up with the output (src:573-609)
	float     lvl_out_raw[kMaxChannels] = { 0, 0 };
	const int rc = phaserot_process_levels (p->dsp, ins, outs, n_samples, angles, lvl_in, lvl_out_raw);
	if (rc != PHASEROT_OK) {
		// real-time context: no way to report; emit silence rather than stale data
		for (uint32_t c = 0; c < p->n_chn; ++c) {
			memset (p->ch[c].out, 0, sizeof (float) * n_samples);
		}
	}
	for (uint32_t c = 0; c < p->n_chn; ++c) {
		// fallback mirror of the angle state, used only if phaserot_plugin_angle() fails
		p->target_state[c] = p->ch[c].last_target;
	}

	// --- meters + notifications -------------------------------------------
	for (uint32_t c = 0; c < p->n_chn; ++c) {
		ChannelState& ch      = p->ch[c];
		lvl_in[c]             = ch.m_in.feed (rc == PHASEROT_OK ? lvl_in[c] : 0.f, p->hold_frames, p->period, p->falloff);
		const float   lvl_out = ch.m_out.feed (rc == PHASEROT_OK ? lvl_out_raw[c] : 0.f, p->hold_frames, p->period, p->falloff);
		float         diff    = 1.0;
		if (ch.m_in.momentary > 0.001f && ch.m_out.momentary > 0.001f) {
			diff        = ch.m_out.momentary / ch.m_in.momentary;
			ch.diff_min = std::fmin (ch.diff_min, diff);
			ch.diff_max = std::fmax (ch.diff_max, diff);
		}
		if (!p->ui_active) {
			continue;
		}
		LV2_Atom_Forge_Frame f;
		lv2_atom_forge_frame_time (&p->forge, 0);
		open_object (p, &f, p->uris.levels);
		lv2_atom_forge_property_head (&p->forge, p->uris.l_channel, 0);
		lv2_atom_forge_int (&p->forge, (int32_t)c);
		put_float (p, p->uris.l_in_cur, lvl_in[c]);
		put_float (p, p->uris.l_in_mom, ch.m_in.momentary);
		put_float (p, p->uris.l_in_peak, ch.m_in.peak);
		put_float (p, p->uris.l_out_cur, lvl_out);
		put_float (p, p->uris.l_out_mom, ch.m_out.momentary);
		put_float (p, p->uris.l_out_peak, ch.m_out.peak);
		put_float (p, p->uris.l_diff_cur, diff);
		put_float (p, p->uris.l_diff_min, ch.diff_min);
		put_float (p, p->uris.l_diff_max, ch.diff_max);
		lv2_atom_forge_pop (&p->forge, &f);
	}
	if (p->ui_active && p->send_state) {
		p->send_state = false;
		send_state (p);
	}
	lv2_atom_forge_pop (&p->forge, &p->seq_frame);
}

const void*
extension_data (const char*)
{
	return nullptr;
}

LV2_Descriptor g_descriptor = { PLUGIN_URI, instantiate, connect_port, activate, run, nullptr, cleanup, extension_data };

} // namespace

extern "C" __attribute__ ((visibility ("default"))) const LV2_Descriptor*
lv2_descriptor (uint32_t index)
{
	switch (index) {
		case 0: g_descriptor.URI = PLUGIN_URI; return &g_descriptor;
		case 1: g_descriptor.URI = PLUGIN_URI "#stereo"; return &g_descriptor;
		default: return nullptr;
	}
}
