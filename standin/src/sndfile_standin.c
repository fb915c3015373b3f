/*
 * Minimal RIFF/WAVE backend for the libsndfile subset in
 * standin/include/sndfile.h.  See that header for scope.
 */
#include <sndfile.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

struct standin_sndfile_s {
	FILE*      fp;
	int        mode;
	SF_INFO    info;
	int        bytes_per_sample;
	int        is_float;
	long       data_offset; /* file offset of first audio byte */
	sf_count_t pos;         /* current frame */
	sf_count_t frames_written;
	unsigned char* scratch;
	size_t     scratch_size;
	const float* mem; /* memory-backed reader (standin_sf_open_memory) */
};

static char g_err[256] = "No Error.";

static uint32_t
rd_u32 (const unsigned char* b)
{
	return (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24);
}
static uint16_t
rd_u16 (const unsigned char* b)
{
	return (uint16_t)(b[0] | (b[1] << 8));
}
static void
wr_u32 (unsigned char* b, uint32_t v)
{
	b[0] = v & 0xff;
	b[1] = (v >> 8) & 0xff;
	b[2] = (v >> 16) & 0xff;
	b[3] = (v >> 24) & 0xff;
}
static void
wr_u16 (unsigned char* b, uint16_t v)
{
	b[0] = v & 0xff;
	b[1] = (v >> 8) & 0xff;
}

static unsigned char*
scratch (SNDFILE* s, size_t n)
{
	if (n > s->scratch_size) {
		free (s->scratch);
		s->scratch      = (unsigned char*)malloc (n);
		s->scratch_size = s->scratch ? n : 0;
	}
	return s->scratch;
}

static int
write_header (SNDFILE* s)
{
	unsigned char h[44];
	const uint32_t bps   = (uint32_t)s->bytes_per_sample;
	const uint32_t ch    = (uint32_t)s->info.channels;
	const uint64_t dlen  = (uint64_t)s->frames_written * bps * ch;
	const uint32_t dlen32 = dlen > 0xffffffffu - 36 ? 0xffffffffu - 36 : (uint32_t)dlen;
	memcpy (h, "RIFF", 4);
	wr_u32 (h + 4, 36 + dlen32);
	memcpy (h + 8, "WAVEfmt ", 8);
	wr_u32 (h + 16, 16);
	wr_u16 (h + 20, s->is_float ? 3 : 1);
	wr_u16 (h + 22, (uint16_t)ch);
	wr_u32 (h + 24, (uint32_t)s->info.samplerate);
	wr_u32 (h + 28, (uint32_t)s->info.samplerate * ch * bps);
	wr_u16 (h + 32, (uint16_t)(ch * bps));
	wr_u16 (h + 34, (uint16_t)(8 * bps));
	memcpy (h + 36, "data", 4);
	wr_u32 (h + 40, dlen32);
	if (fseek (s->fp, 0, SEEK_SET)) {
		return -1;
	}
	return fwrite (h, 1, 44, s->fp) == 44 ? 0 : -1;
}

SNDFILE*
sf_open (const char* path, int mode, SF_INFO* sfinfo)
{
	if (!path || !sfinfo || (mode != SFM_READ && mode != SFM_WRITE)) {
		snprintf (g_err, sizeof (g_err), "Bad parameters to sf_open.");
		return NULL;
	}
	SNDFILE* s = (SNDFILE*)calloc (1, sizeof (*s));
	if (!s) {
		return NULL;
	}
	s->mode = mode;

	if (mode == SFM_WRITE) {
		const int sub = sfinfo->format & SF_FORMAT_SUBMASK;
		if ((sfinfo->format & SF_FORMAT_TYPEMASK) != SF_FORMAT_WAV || sfinfo->channels < 1 || sfinfo->samplerate < 1) {
			snprintf (g_err, sizeof (g_err), "Format not recognised.");
			free (s);
			return NULL;
		}
		switch (sub) {
			case SF_FORMAT_PCM_16: s->bytes_per_sample = 2; break;
			case SF_FORMAT_PCM_24: s->bytes_per_sample = 3; break;
			case SF_FORMAT_PCM_32: s->bytes_per_sample = 4; break;
			case SF_FORMAT_FLOAT: s->bytes_per_sample = 4; s->is_float = 1; break;
			default:
				snprintf (g_err, sizeof (g_err), "Format not recognised.");
				free (s);
				return NULL;
		}
		s->fp = fopen (path, "wb");
		if (!s->fp) {
			snprintf (g_err, sizeof (g_err), "System error : could not open '%s'.", path);
			free (s);
			return NULL;
		}
		s->info        = *sfinfo;
		s->data_offset = 44;
		if (write_header (s)) {
			fclose (s->fp);
			free (s);
			return NULL;
		}
		return s;
	}

	s->fp = fopen (path, "rb");
	if (!s->fp) {
		snprintf (g_err, sizeof (g_err), "System error : could not open '%s'.", path);
		free (s);
		return NULL;
	}
	unsigned char b[40];
	if (fread (b, 1, 12, s->fp) != 12 || memcmp (b, "RIFF", 4) || memcmp (b + 8, "WAVE", 4)) {
		snprintf (g_err, sizeof (g_err), "File contains data in an unknown format.");
		goto fail;
	}
	int      have_fmt = 0;
	uint32_t data_len = 0;
	for (;;) {
		if (fread (b, 1, 8, s->fp) != 8) {
			snprintf (g_err, sizeof (g_err), "No data chunk.");
			goto fail;
		}
		const uint32_t len = rd_u32 (b + 4);
		if (!memcmp (b, "fmt ", 4)) {
			const uint32_t take = len < 40 ? len : 40;
			if (take < 16 || fread (b, 1, take, s->fp) != take) {
				goto fail;
			}
			uint16_t tag        = rd_u16 (b);
			s->info.channels    = rd_u16 (b + 2);
			s->info.samplerate  = (int)rd_u32 (b + 4);
			const uint16_t bits = rd_u16 (b + 14);
			if (tag == 0xFFFE && take >= 26) {
				tag = rd_u16 (b + 24);
			}
			s->bytes_per_sample = bits / 8;
			s->is_float         = (tag == 3);
			if (!((tag == 1 && (bits == 16 || bits == 24 || bits == 32)) || (tag == 3 && bits == 32))) {
				snprintf (g_err, sizeof (g_err), "Unsupported WAV encoding (tag %u, %u bits).", tag, bits);
				goto fail;
			}
			if (len > take) {
				fseek (s->fp, (long)(len - take), SEEK_CUR);
			}
			if (len & 1) {
				fseek (s->fp, 1, SEEK_CUR);
			}
			have_fmt = 1;
		} else if (!memcmp (b, "data", 4)) {
			data_len       = len;
			s->data_offset = ftell (s->fp);
			break;
		} else {
			fseek (s->fp, (long)(len + (len & 1)), SEEK_CUR);
		}
	}
	if (!have_fmt || s->info.channels < 1) {
		snprintf (g_err, sizeof (g_err), "Missing fmt chunk.");
		goto fail;
	}
	{
		/* trust the file size over a clipped/streamed data length */
		long cur = ftell (s->fp);
		fseek (s->fp, 0, SEEK_END);
		long end = ftell (s->fp);
		fseek (s->fp, cur, SEEK_SET);
		uint64_t avail = (uint64_t)(end - cur);
		uint64_t dl    = data_len;
		if (dl > avail || dl >= 0xffffffffu - 36) {
			dl = avail;
		}
		s->info.frames = (sf_count_t)(dl / ((uint64_t)s->bytes_per_sample * s->info.channels));
	}
	s->info.sections = 1;
	s->info.seekable = 1;
	s->info.format   = SF_FORMAT_WAV | (s->is_float ? SF_FORMAT_FLOAT : s->bytes_per_sample == 2 ? SF_FORMAT_PCM_16 : s->bytes_per_sample == 3 ? SF_FORMAT_PCM_24 : SF_FORMAT_PCM_32);
	*sfinfo          = s->info;
	return s;
fail:
	fclose (s->fp);
	free (s);
	return NULL;
}

SNDFILE*
standin_sf_open_memory (const float* interleaved, sf_count_t frames, int channels, int samplerate, SF_INFO* sfinfo)
{
	if (!interleaved || frames < 0 || channels < 1) {
		return NULL;
	}
	SNDFILE* s = (SNDFILE*)calloc (1, sizeof (*s));
	if (!s) {
		return NULL;
	}
	s->mode             = SFM_READ;
	s->mem              = interleaved;
	s->is_float         = 1;
	s->bytes_per_sample = 4;
	s->info.frames      = frames;
	s->info.channels    = channels;
	s->info.samplerate  = samplerate;
	s->info.format      = SF_FORMAT_WAV | SF_FORMAT_FLOAT;
	s->info.sections    = 1;
	s->info.seekable    = 1;
	if (sfinfo) {
		*sfinfo = s->info;
	}
	return s;
}

int
sf_close (SNDFILE* s)
{
	if (!s) {
		return -1;
	}
	if (s->mem) {
		free (s);
		return 0;
	}
	if (s->mode == SFM_WRITE) {
		write_header (s);
	}
	fclose (s->fp);
	free (s->scratch);
	free (s);
	return 0;
}

sf_count_t
sf_readf_float (SNDFILE* s, float* ptr, sf_count_t frames)
{
	if (!s || s->mode != SFM_READ || frames <= 0) {
		return 0;
	}
	if (s->pos + frames > s->info.frames) {
		frames = s->info.frames - s->pos;
	}
	if (frames <= 0) {
		return 0;
	}
	const size_t ns = (size_t)frames * s->info.channels;
	if (s->mem) {
		memcpy (ptr, s->mem + (size_t)s->pos * s->info.channels, ns * sizeof (float));
	} else if (s->is_float) {
		const size_t got = fread (ptr, sizeof (float), ns, s->fp);
		frames           = (sf_count_t)(got / s->info.channels);
	} else {
		unsigned char* b = scratch (s, ns * s->bytes_per_sample);
		if (!b) {
			return 0;
		}
		const size_t got = fread (b, s->bytes_per_sample, ns, s->fp);
		frames           = (sf_count_t)(got / s->info.channels);
		const size_t n   = (size_t)frames * s->info.channels;
		/* libsndfile's default normalisation: integer PCM maps to [-1, 1) */
		if (s->bytes_per_sample == 2) {
			for (size_t i = 0; i < n; ++i) {
				ptr[i] = (float)(int16_t)rd_u16 (b + 2 * i) * (1.0f / 32768.0f);
			}
		} else if (s->bytes_per_sample == 3) {
			for (size_t i = 0; i < n; ++i) {
				int32_t v = (int32_t)((uint32_t)b[3 * i] << 8 | (uint32_t)b[3 * i + 1] << 16 | (uint32_t)b[3 * i + 2] << 24);
				ptr[i]    = (float)v * (1.0f / 2147483648.0f);
			}
		} else {
			for (size_t i = 0; i < n; ++i) {
				ptr[i] = (float)(int32_t)rd_u32 (b + 4 * i) * (1.0f / 2147483648.0f);
			}
		}
	}
	s->pos += frames;
	return frames;
}

/* Integer reads of integer PCM files (libsndfile: no scaling, narrower samples are
 * left-justified in the wider type).  Other sources are outside this subset. */
static sf_count_t
read_pcm (SNDFILE* s, void* ptr, sf_count_t frames, int out_bytes)
{
	if (!s || s->mode != SFM_READ || frames <= 0 || s->mem || s->is_float || s->bytes_per_sample > out_bytes) {
		return 0;
	}
	if (s->pos + frames > s->info.frames) {
		frames = s->info.frames - s->pos;
	}
	if (frames <= 0) {
		return 0;
	}
	const size_t   ns = (size_t)frames * s->info.channels;
	unsigned char* b  = scratch (s, ns * s->bytes_per_sample);
	if (!b) {
		return 0;
	}
	const size_t got = fread (b, s->bytes_per_sample, ns, s->fp);
	frames           = (sf_count_t)(got / s->info.channels);
	const size_t n   = (size_t)frames * s->info.channels;
	for (size_t i = 0; i < n; ++i) {
		uint32_t v = 0; /* left-justified in 32 bits */
		for (int k = 0; k < s->bytes_per_sample; ++k) {
			v |= (uint32_t)b[s->bytes_per_sample * i + k] << (8 * (4 - s->bytes_per_sample + k));
		}
		if (out_bytes == 2) {
			((int16_t*)ptr)[i] = (int16_t)(v >> 16);
		} else {
			((int32_t*)ptr)[i] = (int32_t)v;
		}
	}
	s->pos += frames;
	return frames;
}

/* libsndfile: sf_read_raw copies `bytes` bytes of the data chunk as stored (a whole number of frames). */
sf_count_t
sf_read_raw (SNDFILE* s, void* ptr, sf_count_t bytes)
{
	if (!s || s->mode != SFM_READ || bytes <= 0 || s->mem) {
		return 0;
	}
	const sf_count_t fb = (sf_count_t)s->bytes_per_sample * s->info.channels;
	sf_count_t       frames = bytes / fb;
	if (s->pos + frames > s->info.frames) {
		frames = s->info.frames - s->pos;
	}
	if (frames <= 0) {
		return 0;
	}
	const size_t got = fread (ptr, (size_t)fb, (size_t)frames, s->fp);
	s->pos += (sf_count_t)got;
	return (sf_count_t)got * fb;
}

sf_count_t
sf_readf_short (SNDFILE* s, short* ptr, sf_count_t frames)
{
	return read_pcm (s, ptr, frames, 2);
}

sf_count_t
sf_readf_int (SNDFILE* s, int* ptr, sf_count_t frames)
{
	return read_pcm (s, ptr, frames, 4);
}

static int32_t
clip_scale (float v, double scale, double maxv)
{
	double d = (double)v * scale;
	if (d >= maxv) {
		return (int32_t)maxv;
	}
	if (d <= -maxv - 1.0) {
		return (int32_t)(-maxv - 1.0);
	}
	return (int32_t)lrint (d);
}

sf_count_t
sf_writef_float (SNDFILE* s, const float* ptr, sf_count_t frames)
{
	if (!s || s->mode != SFM_WRITE || frames <= 0) {
		return 0;
	}
	const size_t ns = (size_t)frames * s->info.channels;
	size_t       put;
	if (s->is_float) {
		put = fwrite (ptr, sizeof (float), ns, s->fp);
	} else {
		unsigned char* b = scratch (s, ns * s->bytes_per_sample);
		if (!b) {
			return 0;
		}
		if (s->bytes_per_sample == 2) {
			for (size_t i = 0; i < ns; ++i) {
				wr_u16 (b + 2 * i, (uint16_t)(int16_t)clip_scale (ptr[i], 32768.0, 32767.0));
			}
		} else if (s->bytes_per_sample == 3) {
			for (size_t i = 0; i < ns; ++i) {
				int32_t v     = clip_scale (ptr[i], 8388608.0, 8388607.0);
				b[3 * i]     = v & 0xff;
				b[3 * i + 1] = (v >> 8) & 0xff;
				b[3 * i + 2] = (v >> 16) & 0xff;
			}
		} else {
			for (size_t i = 0; i < ns; ++i) {
				wr_u32 (b + 4 * i, (uint32_t)clip_scale (ptr[i], 2147483648.0, 2147483647.0));
			}
		}
		put = fwrite (b, s->bytes_per_sample, ns, s->fp);
	}
	const sf_count_t done = (sf_count_t)(put / s->info.channels);
	s->frames_written += done;
	return done;
}

sf_count_t
sf_seek (SNDFILE* s, sf_count_t frames, int whence)
{
	if (!s || s->mode != SFM_READ) {
		return -1;
	}
	sf_count_t target;
	switch (whence) {
		case SEEK_SET: target = frames; break;
		case SEEK_CUR: target = s->pos + frames; break;
		case SEEK_END: target = s->info.frames + frames; break;
		default: return -1;
	}
	if (target < 0 || target > s->info.frames) {
		return -1;
	}
	if (!s->mem && fseek (s->fp, s->data_offset + (long)(target * s->bytes_per_sample * s->info.channels), SEEK_SET)) {
		return -1;
	}
	s->pos = target;
	return target;
}

const char*
sf_strerror (SNDFILE* s)
{
	(void)s;
	return g_err;
}

const char*
sf_get_string (SNDFILE* s, int str_type)
{
	(void)s;
	(void)str_type;
	return NULL;
}

int
sf_set_string (SNDFILE* s, int str_type, const char* str)
{
	(void)s;
	(void)str_type;
	(void)str;
	return 0;
}

int
sf_command (SNDFILE* s, int command, void* data, int datasize)
{
	(void)s;
	if (command == SFC_GET_LOG_INFO && data && datasize > 0) {
		snprintf ((char*)data, (size_t)datasize, "standin sndfile: RIFF/WAVE\n");
		return (int)strlen ((char*)data);
	}
	return SF_FALSE;
}
