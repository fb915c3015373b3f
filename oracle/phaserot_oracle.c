/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing in the product path may include,
 * link or execute this file; only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, as the checker.
 *
 * A CPU restatement of the phase-rotation hot path of x42/phaserotate.lv2
 * (reference commit 00fece1).  Every function cites the reference lines it
 * follows (paths relative to /root/reference).
 *
 * How it differs from the reference on purpose: the reference applies its FIR
 * Hilbert transformer by FFT overlap-add convolution in fp32 (FFTW).  A linear
 * convolution does not depend on how it is partitioned, so this restatement
 * evaluates the SAME taps as a direct-form sum in double precision and rounds
 * once to float.  That is the mathematically exact value the reference's FFT
 * path approximates to ~1e-6; everything after the convolution (rotation,
 * peak, block/angle bookkeeping, quirks) follows the reference's fp32
 * arithmetic operation by operation (mul, mul, add — build with
 * -ffp-contract=off).
 *
 * Pinning: the reference ships no golden vectors or tests.  This restatement is
 * pinned against outputs of the reference's own sources, compiled unmodified
 * against stand-in fftw3/sndfile/LV2 headers (oracle/Makefile -> oracle/_ref/),
 * by tests/test_oracle.py, and against fixtures generated from those
 * binaries (tests/golden/, script tests/golden/make_golden.py).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------- */
/* FIR design                                                                */
/* ------------------------------------------------------------------------- */

/*
 * Hilbert FIR taps, L = FIR length (CLI: blksiz, plugin: firlen).
 *
 * Reference: cli/phase-rotate.cc:144-161 and src/phaserotate.c:374-391.
 * Both fill a half spectrum F[k] = (0, +1) for even k, (0, -1) for odd k,
 * k = 0..L/2, take an unnormalised c2r FFT of size L (FFTW ignores the
 * imaginary parts of the DC and Nyquist bins), and multiply each sample by
 * (0.5/L) * (1 - cos(2 pi i / L)) evaluated in double, storing float.
 *
 * The c2r output has the closed form  -2 cot(pi m / L)  for odd m = i - L/2
 * and 0 for even m; it is rounded to float here because the reference holds it
 * in a float array before windowing.
 */
void
pro_fir_taps (int L, float* taps)
{
	const float  normf = 0.5f / (float)L; /* cli:142 `_norm = 0.5 / _parsiz` stored as float */
	const double flen  = 1.0 / (double)L;
	for (int i = 0; i < L; ++i) {
		const int m   = i - L / 2;
		float     raw = 0.f;
		if (m & 1) {
			raw = (float)(-2.0 / tan (M_PI * (double)m / (double)L));
		}
		/* cli:160  fir[i] *= _norm * (1 - cos (2.0 * M_PI * i * flen));
		 * src:387-390 uses the double 0.5/firlen; 0.5/L is exact in float for
		 * every power-of-two L and differs by < 1 ulp(double) otherwise. */
		taps[i] = (float)((double)raw * ((double)normf * (1.0 - cos (2.0 * M_PI * (double)i * flen))));
	}
}

/* Plugin variant: window constant kept in double (src/phaserotate.c:387-391). */
void
pro_fir_taps_plugin (int L, float* taps)
{
	const double fnorm = 0.5 / (double)L;
	const double flen  = 1.0 / (double)L;
	for (int i = 0; i < L; ++i) {
		const int m   = i - L / 2;
		float     raw = 0.f;
		if (m & 1) {
			raw = (float)(-2.0 / tan (M_PI * (double)m / (double)L));
		}
		taps[i] = (float)((double)raw * (fnorm * (1.0 - cos (2.0 * M_PI * (double)i * flen))));
	}
}

/*
 * sin/cos table.  Reference: cli/phase-rotate.cc:38-39, 41-72 with
 * SUBSAMPLE = 2: `float mp = 2.f * M_PI / SUBSAMPLE / -360.0;` (evaluated in
 * double, stored float) and `sincosf (mp * i, ...)` with a float product.
 * Generalised to any subsample factor S; S = 2 is the reference grid.
 * s, c: [180 * S].
 */
void
pro_sincos_lut (int subsample, float* s, float* c)
{
	const float mp = (float)(2.f * M_PI / subsample / -360.0);
	const int   n  = 180 * subsample;
	for (int i = 0; i < n; ++i) {
		const float arg = mp * (float)i;
		s[i]            = sinf (arg);
		c[i]            = cosf (arg);
	}
}

/* ------------------------------------------------------------------------- */
/* Hilbert FIR as an exact linear convolution                                */
/* ------------------------------------------------------------------------- */

/*
 * h[t] = sum_{k<L} taps[k] * x[t - k],  t in [0, n_out), x = 0 outside [0, n).
 *
 * This is what PhaseRotateProc::hilbert's overlap-add (cli/phase-rotate.cc:181-212)
 * and the plugin's partitioned convolution (src/phaserotate.c:633-662) compute,
 * up to fp32 FFT rounding.  Only odd k carry non-zero taps (L/2 is even for
 * every size the reference uses), which halves the work.
 */
typedef struct {
	const float*  x;
	int64_t       n;
	const float*  taps;
	const double* g;
	int           nodd;
	int           even_clean;
	float*        h;
	int64_t       t0, t1;
} fir_job_t;

static void*
fir_worker (void* arg)
{
	const fir_job_t* jb = (const fir_job_t*)arg;
	const int64_t    n  = jb->n;
	for (int64_t t = jb->t0; t < jb->t1; ++t) {
		double acc = 0.0;
		/* odd taps k = 2j+1 read x[t - 1 - 2j]; clip j to where that index is inside [0, n) */
		int64_t j0 = 0;
		int64_t j1 = jb->nodd;
		if (t - 1 - (n - 1) > 0) {
			j0 = (t - 1 - (n - 1) + 1) / 2;
		}
		if (t - 1 < 0) {
			j1 = 0;
		} else if ((t - 1) / 2 + 1 < j1) {
			j1 = (t - 1) / 2 + 1;
		}
		const float* xp = jb->x + (t - 1);
		for (int64_t j = j0; j < j1; ++j) {
			acc += jb->g[j] * (double)xp[-2 * j];
		}
		if (!jb->even_clean) {
			for (int j = 0; j < jb->nodd; ++j) {
				const int64_t idx = t - 2 * (int64_t)j;
				if (idx >= 0 && idx < n) {
					acc += (double)jb->taps[2 * j] * (double)jb->x[idx];
				}
			}
		}
		jb->h[t] = (float)acc;
	}
	return NULL;
}

void
pro_hilbert_fir (const float* x, int64_t n, const float* taps, int L, float* h, int64_t n_out)
{
	const int nodd = L / 2;
	double*   g    = (double*)malloc (sizeof (double) * (size_t)nodd);
	int       even_clean = 1;
	for (int j = 0; j < nodd; ++j) {
		g[j] = (double)taps[2 * j + 1];
		if (taps[2 * j] != 0.f) {
			even_clean = 0;
		}
	}
	long nthr = sysconf (_SC_NPROCESSORS_ONLN);
	if (nthr < 1) nthr = 1;
	if (nthr > 64) nthr = 64;
	if (n_out < 4096) nthr = 1;
	pthread_t th[64];
	fir_job_t jobs[64];
	for (long i = 0; i < nthr; ++i) {
		jobs[i] = (fir_job_t){ x, n, taps, g, nodd, even_clean, h, n_out * i / nthr, n_out * (i + 1) / nthr };
		if (nthr == 1) {
			fir_worker (&jobs[i]);
		} else {
			pthread_create (&th[i], NULL, fir_worker, &jobs[i]);
		}
	}
	if (nthr > 1) {
		for (long i = 0; i < nthr; ++i) {
			pthread_join (th[i], NULL);
		}
	}
	free (g);
}

/* ------------------------------------------------------------------------- */
/* CLI analysis                                                              */
/* ------------------------------------------------------------------------- */

/* cli/dsp_peak_calc.h:314-324 (scalar form; the SIMD variants compute the same
 * order-independent maximum of |x|). */
static float
peak_abs (const float* buf, int64_t n, float current)
{
	for (int64_t i = 0; i < n; ++i) {
		const float v = fabsf (buf[i]);
		if (v > current) {
			current = v;
		}
	}
	return current;
}

/* cli/phase-rotate.cc:98-121: x = ca*b0[i] + sa*b1[i] (two rounded products,
 * one rounded sum), running max of |x|. */
static float
rotated_peak (const float* b0, const float* b1, int64_t n, float pk, float sa, float ca)
{
	for (int64_t i = 0; i < n; ++i) {
		const float p0 = ca * b0[i];
		const float p1 = sa * b1[i];
		const float v  = fabsf (p0 + p1);
		if (v > pk) {
			pk = v;
		}
	}
	return pk;
}

/*
 * One whole-file analysis pass: analyze_file (cli/phase-rotate.cc:565-587)
 * driving PhaseRotate::analyze / thr_process (cli:388-444).
 *
 *   interleaved : n_frames * n_chn floats
 *   subsample   : angle grid, MAXSAMPLE = 180 * subsample (reference: 2)
 *   peaks       : [n_chn][MAXSAMPLE], read-modify-write like PhaseRotate::_peak
 *                 (callers zero it to model PhaseRotate::reset, cli:355-366)
 *
 * Block structure: B = ceil(F / L) real blocks (the short last one zero padded,
 * cli:577-580) followed by one all-zero flush block (cli:585-586).  In block n
 * (time base t0 = n L) thr_process sees
 *     tdc = [ x[t0-L .. t0) | x[t0 .. t0+L) ],  hil[i] = H[t0 + i].
 * Angle loop (cli:409-428): `angle` runs from ang_start in steps of ang_stride
 * while angle <= ang_end, leaving as soon as angle >= ang_end after the step.
 *   - un-wrapped angle == 0: raw input peak of the new block (cli:413-414);
 *   - first block (start): only i in [L/2, L) is examined, against tdc[L/2..L),
 *     i.e. against all-zero history (cli:418-419);
 *   - otherwise max_i |ca * x[t0 + i - L/2] + sa * H[t0 + i]|, i in [0, L) (cli:421).
 */
/*
 * Shard form used by the multi-GPU tests: the stream continues a longer one.
 *   hist  : blksiz frames (interleaved) that precede `interleaved`, or NULL (silence)
 *   first : the shard starts the stream -> first-block rule applies to its block 0
 *   last  : the shard ends the stream  -> short block zero padded + zero flush block;
 *           otherwise n_frames must be a multiple of blksiz and no flush block is run
 * pro_cli_analyze() is the whole-file case (hist NULL, first = last = 1).
 */
void
pro_cli_analyze_shard (const float* interleaved, int64_t n_frames, int n_chn, int blksiz, int subsample,
                       const float* hist, int first, int last,
                       int ang_start, int ang_end, int ang_stride, int only_chn, float* peaks)
{
	const int     L         = blksiz;
	const int     D         = L / 2;
	const int     maxsample = 180 * subsample;
	const int64_t B         = last ? (n_frames + L - 1) / L : n_frames / L;
	const int64_t n_blocks  = last ? B + 1 : B; /* real blocks (+ flush block) */
	const int64_t n_pad     = n_blocks * (int64_t)L;

	float* lut_s = (float*)malloc (sizeof (float) * (size_t)maxsample);
	float* lut_c = (float*)malloc (sizeof (float) * (size_t)maxsample);
	float* taps  = (float*)malloc (sizeof (float) * (size_t)L);
	pro_sincos_lut (subsample, lut_s, lut_c);
	pro_fir_taps (L, taps);

	/* x with one block of history in front: xz[L + t] = x[t] */
	float* xz = (float*)calloc ((size_t)(n_pad + L), sizeof (float));
	float* H  = (float*)malloc (sizeof (float) * (size_t)(n_pad + L));

	const int c0 = only_chn < 0 ? 0 : only_chn;
	const int c1 = only_chn < 0 ? n_chn : only_chn + 1;

	for (int c = c0; c < c1; ++c) {
		for (int i = 0; i < L; ++i) {
			xz[i] = hist ? hist[(size_t)i * n_chn + c] : 0.f;
		}
		for (int64_t t = 0; t < n_frames; ++t) {
			xz[L + t] = interleaved[t * n_chn + c];
		}
		/* convolve from the start of the history so that H over the shard sees it;
		 * H[L + t] is the value at shard time t */
		pro_hilbert_fir (xz, n_frames + L, taps, L, H, n_pad + L);
		float* pk = peaks + (size_t)c * maxsample;

		for (int64_t n = 0; n < n_blocks; ++n) {
			const int     start = (first && n == 0 && B > 0);
			const float*  blk   = xz + L + n * (int64_t)L; /* new block */
			const float*  dly   = blk - D;                /* &tdc[firlen] */
			const float*  hil   = H + L + n * (int64_t)L;
			int           angle = ang_start;
			while (angle <= ang_end) {
				const int a = ((angle % maxsample) + maxsample) % maxsample;
				if (angle == 0) {
					pk[a] = peak_abs (blk, L, pk[a]);
				} else if (start) {
					/* history is all zero here (start of the stream) */
					pk[a] = rotated_peak (dly, hil + D, D, pk[a], lut_s[a], lut_c[a]);
				} else {
					pk[a] = rotated_peak (dly, hil, L, pk[a], lut_s[a], lut_c[a]);
				}
				angle += ang_stride;
				if (angle >= ang_end) {
					break;
				}
			}
		}
	}
	free (H);
	free (xz);
	free (taps);
	free (lut_c);
	free (lut_s);
}

void
pro_cli_analyze (const float* interleaved, int64_t n_frames, int n_chn, int blksiz, int subsample,
                 int ang_start, int ang_end, int ang_stride, int only_chn, float* peaks)
{
	pro_cli_analyze_shard (interleaved, n_frames, n_chn, blksiz, subsample, NULL, 1, 1, ang_start, ang_end, ang_stride, only_chn, peaks);
}

/*
 * Oversampled true-peak variant of the analysis pass.  NOT a reference feature:
 * the reference measures the digital peak only (cli:98-121), so there is
 * nothing to pin this against - "parity unpinned".  It restates the definition
 * in include/phaserot_cuda.h so that the GPU path has an independent checker:
 *
 *   interpolator  ITU-R BS.1770-4 Annex 2, 48 taps in 4 phases of 12,
 *                 s^[t, ph] = sum_k c[ph][k] s[t - k]   (oversample 2: phases 0, 2)
 *   pair          p[t] = (xq[t], H[t]),  xq[t] = x[t - L/2], forced to 0 for
 *                 t < L when the first-block rule applies (cli:418-419)
 *   peak[c][a]    max over the samples the reference examines (block / angle
 *                 bookkeeping exactly as pro_cli_analyze_shard) of
 *                 max (|ca xq + sa H|, max_ph |ca xq^ + sa H^|)
 *   un-wrapped angle 0 (cli:413-414): the same detector on the raw input block.
 */
static const float tp_c0[12] = { 0.0017089843750f, 0.0109863281250f, -0.0196533203125f, 0.0332031250000f, -0.0594482421875f, 0.1373291015625f,
	                             0.9721679687500f, -0.1022949218750f, 0.0476074218750f, -0.0266113281250f, 0.0148925781250f, -0.0083007812500f };
static const float tp_c1[12] = { -0.0291748046875f, 0.0292968750000f, -0.0517578125000f, 0.0891113281250f, -0.1665039062500f, 0.4650878906250f,
	                             0.7797851562500f, -0.2003173828125f, 0.1015625000000f, -0.0582275390625f, 0.0330810546875f, -0.0189208984375f };

static float
tp_coef (int ph, int k)
{
	return ph == 0 ? tp_c0[k] : ph == 1 ? tp_c1[k] : ph == 2 ? tp_c1[11 - k] : tp_c0[11 - k];
}

/* s points at time 0 of an array that is readable from index -11 */
static float
tp_interp (const float* s, int64_t t, int ph)
{
	float a = 0.f;
	for (int k = 0; k < 12; ++k) {
		a = fmaf (tp_coef (ph, k), s[t - k], a);
	}
	return a;
}

void
pro_cli_analyze_tp_shard (const float* interleaved, int64_t n_frames, int n_chn, int blksiz, int subsample, int oversample,
                          const float* hist, int first, int last,
                          int ang_start, int ang_end, int ang_stride, int only_chn, float* peaks)
{
	const int     L         = blksiz;
	const int     D         = L / 2;
	const int     maxsample = 180 * subsample;
	const int64_t B         = last ? (n_frames + L - 1) / L : n_frames / L;
	const int64_t n_blocks  = last ? B + 1 : B;
	const int64_t n_pad     = n_blocks * (int64_t)L;
	const int     nph       = oversample == 4 ? 4 : 2;
	const int     phs[4]    = { 0, oversample == 4 ? 1 : 2, 2, 3 };

	float* lut_s = (float*)malloc (sizeof (float) * (size_t)maxsample);
	float* lut_c = (float*)malloc (sizeof (float) * (size_t)maxsample);
	float* taps  = (float*)malloc (sizeof (float) * (size_t)L);
	pro_sincos_lut (subsample, lut_s, lut_c);
	pro_fir_taps (L, taps);

	/* arrays indexed [L + t], t = shard time; L frames of history in front */
	float* xz = (float*)calloc ((size_t)(n_pad + L), sizeof (float)); /* raw input */
	float* xq = (float*)calloc ((size_t)(n_pad + L), sizeof (float)); /* delayed direct branch incl. first-block rule */
	float* H  = (float*)malloc (sizeof (float) * (size_t)(n_pad + L));
	/* interpolated [ph][t], t in [0, n_pad) */
	float* xi = (float*)malloc (sizeof (float) * (size_t)n_pad * 4);
	float* qi = (float*)malloc (sizeof (float) * (size_t)n_pad * 4);
	float* hi = (float*)malloc (sizeof (float) * (size_t)n_pad * 4);

	const int c0 = only_chn < 0 ? 0 : only_chn;
	const int c1 = only_chn < 0 ? n_chn : only_chn + 1;

	for (int c = c0; c < c1; ++c) {
		memset (xz, 0, sizeof (float) * (size_t)(n_pad + L));
		for (int i = 0; i < L; ++i) {
			xz[i] = hist ? hist[(size_t)i * n_chn + c] : 0.f;
		}
		for (int64_t t = 0; t < n_frames; ++t) {
			xz[L + t] = interleaved[t * n_chn + c];
		}
		pro_hilbert_fir (xz, n_frames + L, taps, L, H, n_pad + L);
		/* cli:418-419: in the first block of a stream the cos term sees zero history */
		const int64_t t_zero = (first && B > 0) ? L : -(int64_t)L;
		for (int64_t t = -(int64_t)L; t < n_pad; ++t) {
			xq[L + t] = (t >= t_zero && t - D >= -(int64_t)L) ? xz[L + t - D] : 0.f;
		}
		for (int p = 0; p < nph; ++p) {
			for (int64_t t = 0; t < n_pad; ++t) {
				xi[(size_t)p * n_pad + t] = tp_interp (xz + L, t, phs[p]);
				qi[(size_t)p * n_pad + t] = tp_interp (xq + L, t, phs[p]);
				hi[(size_t)p * n_pad + t] = tp_interp (H + L, t, phs[p]);
			}
		}
		float* pk = peaks + (size_t)c * maxsample;

		for (int64_t n = 0; n < n_blocks; ++n) {
			const int     start = (first && n == 0 && B > 0);
			const int64_t tb    = n * (int64_t)L;
			int           angle = ang_start;
			while (angle <= ang_end) {
				const int a = ((angle % maxsample) + maxsample) % maxsample;
				if (angle == 0) {
					pk[a] = peak_abs (xz + L + tb, L, pk[a]);
					for (int p = 0; p < nph; ++p) {
						pk[a] = peak_abs (xi + (size_t)p * n_pad + tb, L, pk[a]);
					}
				} else {
					const int64_t o = start ? D : 0, len = start ? D : L;
					pk[a] = rotated_peak (xq + L + tb + o, H + L + tb + o, len, pk[a], lut_s[a], lut_c[a]);
					for (int p = 0; p < nph; ++p) {
						pk[a] = rotated_peak (qi + (size_t)p * n_pad + tb + o, hi + (size_t)p * n_pad + tb + o, len, pk[a], lut_s[a], lut_c[a]);
					}
				}
				angle += ang_stride;
				if (angle >= ang_end) {
					break;
				}
			}
		}
	}
	free (hi);
	free (qi);
	free (xi);
	free (H);
	free (xq);
	free (xz);
	free (taps);
	free (lut_c);
	free (lut_s);
}

/* ------------------------------------------------------------------------- */
/* CLI render                                                                */
/* ------------------------------------------------------------------------- */

/*
 * Block stream through PhaseRotate::apply (cli/phase-rotate.cc:446-485) =
 * hilbert + rotate (cli:214-232): per channel, for every t >= 0,
 *     y[t] = ca * x[t - L/2] + sa * H[t]
 * over ceil(F/L) zero-padded blocks plus `flush_blocks` zero blocks.
 * angles[c] are half-degree style indices, wrapped like cli:463.
 * out: interleaved, (ceil(F/L) + flush_blocks) * L frames.
 */
void
pro_cli_apply (const float* interleaved, int64_t n_frames, int n_chn, int blksiz, int subsample,
               const int* angles, int flush_blocks, float* out)
{
	const int     L         = blksiz;
	const int     D         = L / 2;
	const int     maxsample = 180 * subsample;
	const int64_t B         = (n_frames + L - 1) / L;
	const int64_t n_out     = (B + flush_blocks) * (int64_t)L;

	float* lut_s = (float*)malloc (sizeof (float) * (size_t)maxsample);
	float* lut_c = (float*)malloc (sizeof (float) * (size_t)maxsample);
	float* taps  = (float*)malloc (sizeof (float) * (size_t)L);
	pro_sincos_lut (subsample, lut_s, lut_c);
	pro_fir_taps (L, taps);

	float* xz = (float*)calloc ((size_t)(n_out + L), sizeof (float));
	float* H  = (float*)malloc (sizeof (float) * (size_t)n_out);
	for (int c = 0; c < n_chn; ++c) {
		memset (xz, 0, sizeof (float) * (size_t)(n_out + L));
		for (int64_t t = 0; t < n_frames && t < n_out; ++t) {
			xz[L + t] = interleaved[t * n_chn + c];
		}
		pro_hilbert_fir (xz + L, n_frames, taps, L, H, n_out);
		const int   a  = ((angles[c] % maxsample) + maxsample) % maxsample;
		const float sa = lut_s[a];
		const float ca = lut_c[a];
		for (int64_t t = 0; t < n_out; ++t) {
			const float p0       = ca * xz[L + t - D];
			const float p1       = sa * H[t];
			out[t * n_chn + c] = p0 + p1;
		}
	}
	free (H);
	free (xz);
	free (taps);
	free (lut_c);
	free (lut_s);
}

/*
 * File render loop of main() (cli/phase-rotate.cc:950-1003) including the
 * latency trim and its two quirks, applied on top of pro_cli_apply's stream:
 *   R1: the first write starts at FLOAT offset `latency` into the interleaved
 *       buffer (`&buf[off]`, cli:985), i.e. at frame latency / n_chn, channel
 *       latency % n_chn, and still writes blksiz - latency frames.
 *   R2: a short last block with n >= latency is not zero filled (cli:973): the
 *       frames beyond n still hold the previous block's *rendered* output.
 * out receives exactly the frames the reference writes; returns their count.
 * `out` must hold n_frames + blksiz frames.
 */
int64_t
pro_cli_render_file (const float* interleaved, int64_t n_frames, int n_chn, int blksiz, int subsample,
                     const int* angles, float* out)
{
	const int      L       = blksiz;
	const uint32_t latency = (uint32_t)L / 2;
	const int      D       = L / 2;
	const int      maxsample = 180 * subsample;

	float* lut_s = (float*)malloc (sizeof (float) * (size_t)maxsample);
	float* lut_c = (float*)malloc (sizeof (float) * (size_t)maxsample);
	float* taps  = (float*)malloc (sizeof (float) * (size_t)L);
	pro_sincos_lut (subsample, lut_s, lut_c);
	pro_fir_taps (L, taps);

	/* The stale-tail quirk feeds rendered output back in as input, so the
	 * stream has to be simulated block by block: state = previous input block
	 * (history) per channel, and H needs the previous and current block. */
	const size_t bs   = (size_t)L * n_chn;
	float*       buf  = (float*)calloc (bs + (size_t)L, sizeof (float)); /* slack: R1 reads past the end never, but keep margin */
	float*       prev = (float*)calloc (bs, sizeof (float)); /* de-interleaved history [c][L] */
	float*       cur  = (float*)calloc (bs, sizeof (float)); /* de-interleaved new block */
	float*       xx   = (float*)calloc ((size_t)3 * L, sizeof (float));
	float*       hh   = (float*)calloc ((size_t)L, sizeof (float));
	/* overlap-add state is equivalent to knowing the previous input block */

	int64_t  written = 0;
	int64_t  pos     = 0;
	uint32_t pad     = 0;
	uint32_t off     = latency;

	for (;;) {
		/* n = sf_readf_float (infile, buf, blksiz); if (n <= 0) break;  (cli:968-971) */
		int64_t n = n_frames - pos;
		if (n > L) {
			n = L;
		}
		if (n <= 0) {
			break;
		}
		memcpy (buf, interleaved + (size_t)pos * n_chn, sizeof (float) * (size_t)n * n_chn);
		pos += n;
		if ((uint32_t)n < latency) {
			pad = (uint32_t)L - (uint32_t)n;
			memset (&buf[n_chn * n], 0, sizeof (float) * (size_t)n_chn * pad);
			pad = latency - (uint32_t)n;
			n += pad;
		}
		/* pr.apply (buf, angles) */
		for (int c = 0; c < n_chn; ++c) {
			float* pc = prev + (size_t)c * L;
			float* cc = cur + (size_t)c * L;
			for (int i = 0; i < L; ++i) {
				cc[i] = buf[c + i * n_chn];
			}
			/* xx = [prev | cur | 0]; H over the cur block only needs prev+cur */
			memcpy (xx, pc, sizeof (float) * (size_t)L);
			memcpy (xx + L, cc, sizeof (float) * (size_t)L);
			/* h[i] = sum_k taps[k] xx[L + i - k] */
			for (int i = 0; i < L; ++i) {
				double acc = 0.0;
				for (int k = 1; k < L; k += 2) {
					acc += (double)taps[k] * (double)xx[L + i - k];
				}
				hh[i] = (float)acc;
			}
			const int   a  = ((angles[c] % maxsample) + maxsample) % maxsample;
			const float sa = lut_s[a];
			const float ca = lut_c[a];
			for (int i = 0; i < L; ++i) {
				const float p0      = ca * xx[L + i - D];
				const float p1      = sa * hh[i];
				buf[c + i * n_chn] = p0 + p1;
			}
			memcpy (pc, cc, sizeof (float) * (size_t)L);
		}
		n -= off;
		memcpy (out + (size_t)written * n_chn, &buf[off], sizeof (float) * (size_t)n * n_chn);
		written += n;
		off = 0;
	}

	const int64_t nflush = (int64_t)latency - pad;
	if (nflush > 0) {
		memset (buf, 0, bs * sizeof (float));
		for (int c = 0; c < n_chn; ++c) {
			float* pc = prev + (size_t)c * L;
			memcpy (xx, pc, sizeof (float) * (size_t)L);
			memset (xx + L, 0, sizeof (float) * (size_t)L);
			for (int i = 0; i < L; ++i) {
				double acc = 0.0;
				for (int k = 1; k < L; k += 2) {
					acc += (double)taps[k] * (double)xx[L + i - k];
				}
				hh[i] = (float)acc;
			}
			const int   a  = ((angles[c] % maxsample) + maxsample) % maxsample;
			const float sa = lut_s[a];
			const float ca = lut_c[a];
			for (int i = 0; i < L; ++i) {
				const float p0      = ca * xx[L + i - D];
				const float p1      = sa * hh[i];
				buf[c + i * n_chn] = p0 + p1;
			}
		}
		memcpy (out + (size_t)written * n_chn, buf, sizeof (float) * (size_t)nflush * n_chn);
		written += nflush;
	}

	free (hh);
	free (xx);
	free (cur);
	free (prev);
	free (buf);
	free (taps);
	free (lut_c);
	free (lut_s);
	return written;
}

/* ------------------------------------------------------------------------- */
/* Plugin run()                                                              */
/* ------------------------------------------------------------------------- */

/* src/phaserotate.c:278-297 */
void
pro_plugin_sizes (double rate, uint32_t* fftlen, uint32_t* firlen, uint32_t* parsiz, uint32_t* latency)
{
	uint32_t fl, fir;
	if (rate < 64000) {
		fl  = 512;
		fir = 3072;
	} else if (rate < 128000) {
		fl  = 1024;
		fir = 4096;
	} else {
		fl  = 2048;
		fir = 8192;
	}
	if (fftlen) *fftlen = fl;
	if (firlen) *firlen = fir;
	if (parsiz) *parsiz = fl / 2;
	if (latency) *latency = fl / 2 + fir / 2;
}

/* src/phaserotate.c:122-133 */
static void
plugin_sin_cos (float angle, float* s, float* c)
{
	static const float twopi = (float)(2 * M_PI);
	const float        a     = angle * twopi;
	*s = sinf (a);
	*c = cosf (a);
}

/*
 * One channel of the plugin's audio path: run() -> process_channel()
 * (src/phaserotate.c:538-725) after activate() on a fresh instance
 * (angle state starts at 0, src:147; buffers zero, src:169-177).
 *
 *   in, out  : n_frames mono samples
 *   block    : frames per run() call (the last call may be shorter)
 *   angles   : port value in degrees for each call, [ceil(n_frames / block)]
 *
 * Closed form used: partition n (P = parsiz frames) is processed
 * when its last input frame arrives and emitted during the next P frames,
 *   out[(n+1)P + i] = ca * x[nP + i - firlat] + sa * sum_k fir[k] x[nP + i - k],
 * out[0..P) = 0.  (ca, sa) follow the ramp of src:673-717 using the target
 * angle of the run() call in which the partition completes (src:564-571).
 */
void
pro_plugin_run (double rate, const float* in, float* out, int64_t n_frames, uint32_t block, const float* angles)
{
	uint32_t fftlen, firlen, P, latency;
	pro_plugin_sizes (rate, &fftlen, &firlen, &P, &latency);
	const uint32_t firlat    = firlen / 2;
	const float    interp_th = (float)P * 1e-6f; /* src:295 */
	const float    interp_nm = 1.f / (float)P;   /* src:296 */

	float* taps = (float*)malloc (sizeof (float) * firlen);
	pro_fir_taps_plugin ((int)firlen, taps);

	float* H = (float*)malloc (sizeof (float) * (size_t)(n_frames > 0 ? n_frames : 1));
	pro_hilbert_fir (in, n_frames, taps, (int)firlen, H, n_frames);

	float angle = 0.f; /* channel_init, src:147 */
	float sa, ca;
	plugin_sin_cos (angle, &sa, &ca); /* src:159 */

	for (int64_t t = 0; t < n_frames && t < (int64_t)P; ++t) {
		out[t] = 0.f;
	}

	const int64_t n_part = n_frames / P; /* partitions that complete */
	for (int64_t n = 0; n < n_part; ++n) {
		/* the call during which input frame (n+1)P - 1 arrives */
		const int64_t call   = ((n + 1) * (int64_t)P - 1) / block;
		float         target = angles[call] / -360.f; /* src:564 */
		if (target < -.5f) target = -.5f;
		if (target > 0.5f) target = 0.5f;

		const int64_t base = n * (int64_t)P;
		float         y[2048];
		if (target != angle) {
			float da = target - angle; /* src:675 */
			if (fabs (da) > 0.5) {
				if (da < 0) {
					da += 1.f;
				} else {
					da -= 1.f;
				}
			}
			da *= interp_nm;
			int fin = 0;
			if (da > interp_th) {
				da = interp_th;
			} else if (da < -interp_th) {
				da = -interp_th;
			} else {
				fin = 1;
			}
			for (uint32_t i = 0; i < P; ++i) {
				float c_, s_;
				plugin_sin_cos (angle, &s_, &c_);
				const int64_t ti = base + i - firlat;
				const float   xd = ti >= 0 ? in[ti] : 0.f;
				const float   p0 = c_ * xd;
				const float   p1 = s_ * H[base + i];
				y[i]             = p0 + p1;
				angle += da;
			}
			if (fin) {
				angle = target;
			}
			if (angle == target) {
				plugin_sin_cos (angle, &sa, &ca);
			}
		} else {
			for (uint32_t i = 0; i < P; ++i) {
				const int64_t ti = base + i - firlat;
				const float   xd = ti >= 0 ? in[ti] : 0.f;
				const float   p0 = ca * xd;
				const float   p1 = sa * H[base + i];
				y[i]             = p0 + p1;
			}
		}
		for (uint32_t i = 0; i < P; ++i) {
			const int64_t to = base + P + i;
			if (to < n_frames) {
				out[to] = y[i];
			}
		}
	}
	free (H);
	free (taps);
}
