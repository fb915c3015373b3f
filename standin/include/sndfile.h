/*
 * Stand-in for <sndfile.h>: libsndfile is not installed in this image.
 *
 * Declares the subset of the libsndfile API used by the reference CLI
 * (cli/phase-rotate.cc:541-563, 573, 685-702, 710, 872, 955, 968, 985, 998,
 * 1002, 1007) and by this repository's own host CLI, with libsndfile's
 * numeric constants.  standin/src/sndfile_standin.c implements it for
 * RIFF/WAVE files (PCM 16/24/32 and IEEE float 32) only; string metadata,
 * cue points and broadcast info are reported as absent.
 *
 * A deployment links the real libsndfile instead; nothing else changes.
 */
#ifndef STANDIN_SNDFILE_H
#define STANDIN_SNDFILE_H

#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int64_t sf_count_t;
typedef struct standin_sndfile_s SNDFILE;

typedef struct SF_INFO {
	sf_count_t frames;
	int        samplerate;
	int        channels;
	int        format;
	int        sections;
	int        seekable;
} SF_INFO;

enum {
	SF_FORMAT_WAV    = 0x010000,
	SF_FORMAT_PCM_16 = 0x0002,
	SF_FORMAT_PCM_24 = 0x0003,
	SF_FORMAT_PCM_32 = 0x0004,
	SF_FORMAT_FLOAT  = 0x0006,

	SF_FORMAT_SUBMASK  = 0x0000FFFF,
	SF_FORMAT_TYPEMASK = 0x0FFF0000
};

enum {
	SF_FALSE = 0,
	SF_TRUE  = 1,

	SFM_READ  = 0x10,
	SFM_WRITE = 0x20,
	SFM_RDWR  = 0x30
};

enum {
	SFC_GET_LOG_INFO       = 0x1001,
	SFC_GET_CUE            = 0x10CE,
	SFC_SET_CUE            = 0x10CF,
	SFC_GET_BROADCAST_INFO = 0x10F0,
	SFC_SET_BROADCAST_INFO = 0x10F1
};

enum {
	SF_STR_TITLE = 0x01,
	SF_STR_GENRE = 0x10
};
#define SF_STR_FIRST SF_STR_TITLE
#define SF_STR_LAST SF_STR_GENRE

typedef struct {
	int32_t  indx;
	uint32_t position;
	int32_t  fcc_chunk;
	int32_t  chunk_start;
	int32_t  block_start;
	uint32_t sample_offset;
	char     name[256];
} SF_CUE_POINT;

typedef struct {
	uint32_t     cue_count;
	SF_CUE_POINT cue_points[100];
} SF_CUES;

typedef struct {
	char     description[256];
	char     originator[32];
	char     originator_reference[32];
	char     origination_date[10];
	char     origination_time[8];
	uint32_t time_reference_low;
	uint32_t time_reference_high;
	short    version;
	char     umid[64];
	char     reserved[190];
	uint32_t coding_history_size;
	char     coding_history[256];
} SF_BROADCAST_INFO;

SNDFILE*    sf_open (const char* path, int mode, SF_INFO* sfinfo);
int         sf_close (SNDFILE* sndfile);
sf_count_t  sf_readf_float (SNDFILE* sndfile, float* ptr, sf_count_t frames);
sf_count_t  sf_readf_short (SNDFILE* sndfile, short* ptr, sf_count_t frames); /* 16-bit PCM files: the samples as stored */
sf_count_t  sf_readf_int (SNDFILE* sndfile, int* ptr, sf_count_t frames);     /* integer PCM files: samples left-justified in 32 bits */
sf_count_t  sf_read_raw (SNDFILE* sndfile, void* ptr, sf_count_t bytes);          /* the data chunk as stored */
sf_count_t  sf_writef_float (SNDFILE* sndfile, const float* ptr, sf_count_t frames);
sf_count_t  sf_seek (SNDFILE* sndfile, sf_count_t frames, int whence);
const char* sf_strerror (SNDFILE* sndfile);
const char* sf_get_string (SNDFILE* sndfile, int str_type);
int         sf_set_string (SNDFILE* sndfile, int str_type, const char* str);
int         sf_command (SNDFILE* sndfile, int command, void* data, int datasize);

/* Stand-in extension (not part of libsndfile): read-only SNDFILE over an
 * interleaved float array that the caller keeps alive.  Lets a test harness
 * hand in-memory audio to code written against the libsndfile read API. */
SNDFILE* standin_sf_open_memory (const float* interleaved, sf_count_t frames, int channels, int samplerate, SF_INFO* sfinfo);

#ifdef __cplusplus
}
#endif
#endif
