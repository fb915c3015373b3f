/*
 * Stand-in implementation of the fftwf_* subset declared in
 * standin/include/fftw3.h.  Test / oracle infrastructure only.
 *
 * Conventions reproduced (FFTW3 manual, "What FFTW Really Computes"):
 *   r2c:  X[k] = sum_j x[j] exp(-2 pi i j k / n),  k = 0..n/2   (unnormalised)
 *   c2r:  x[j] = sum_k X[k] exp(+2 pi i j k / n) over the Hermitian extension
 *         of X[0..n/2]; the imaginary parts of X[0] and X[n/2] do not
 *         contribute.  c2r(r2c(x)) = n * x.
 *
 * Arithmetic: inputs are widened to `real_t`, the transform runs in `real_t`
 * and the result is rounded once to float.  real_t is double by default, which
 * makes this a correctly-rounded DFT for the sizes used (the parity oracle);
 * -DSTANDIN_FFT_FLOAT selects float, which is what the timed CPU baseline uses
 * so that the baseline is not slowed down by double arithmetic.
 *
 * Power-of-two n uses a half-size complex radix-2 FFT plus the usual real
 * split step; any other even/odd n (the plugin designs its FIR with n = 3072,
 * src/phaserotate.c:364) falls back to a table-driven O(n^2) DFT.
 */
#define _POSIX_C_SOURCE 200112L
#include <fftw3.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef STANDIN_FFT_FLOAT
typedef float real_t;
#else
typedef double real_t;
#endif

struct standin_fftwf_plan_s {
	int      n;      /* logical (real) transform size */
	int      pow2;   /* n is a power of two >= 4 */
	int      c2r;    /* direction */
	int      m;      /* n / 2 : size of the inner complex FFT */
	int      logm;
	real_t*  tw_re;  /* exp(-2 pi i k / m), k < m/2  (inner FFT twiddles) */
	real_t*  tw_im;
	real_t*  sp_re;  /* exp(-2 pi i k / n), k <= m   (split step) / k < n (naive) */
	real_t*  sp_im;
	uint32_t* rev;   /* bit reversal of 0..m-1 */
	real_t*  wr;     /* work: m complex */
	real_t*  wi;
};

void*
fftwf_malloc (size_t n)
{
	void* p = NULL;
	if (posix_memalign (&p, 64, n ? n : 1)) {
		return NULL;
	}
	return p;
}

void
fftwf_free (void* p)
{
	free (p);
}

static fftwf_plan
plan_new (int n, int c2r)
{
	if (n < 1) {
		return NULL;
	}
	fftwf_plan p = (fftwf_plan)calloc (1, sizeof (*p));
	if (!p) {
		return NULL;
	}
	p->n   = n;
	p->c2r = c2r;
	p->pow2 = (n >= 4) && ((n & (n - 1)) == 0);

	const double twopi = 6.283185307179586476925286766559;

	if (p->pow2) {
		p->m = n / 2;
		for (p->logm = 0; (1 << p->logm) < p->m; ++p->logm)
			;
		p->tw_re = (real_t*)malloc (sizeof (real_t) * (p->m / 2 + 1));
		p->tw_im = (real_t*)malloc (sizeof (real_t) * (p->m / 2 + 1));
		p->sp_re = (real_t*)malloc (sizeof (real_t) * (p->m + 1));
		p->sp_im = (real_t*)malloc (sizeof (real_t) * (p->m + 1));
		p->rev   = (uint32_t*)malloc (sizeof (uint32_t) * p->m);
		p->wr    = (real_t*)malloc (sizeof (real_t) * p->m);
		p->wi    = (real_t*)malloc (sizeof (real_t) * p->m);
		for (int k = 0; k < p->m / 2 + 1; ++k) {
			p->tw_re[k] = (real_t)cos (twopi * k / p->m);
			p->tw_im[k] = (real_t)-sin (twopi * k / p->m);
		}
		for (int k = 0; k <= p->m; ++k) {
			p->sp_re[k] = (real_t)cos (twopi * k / n);
			p->sp_im[k] = (real_t)-sin (twopi * k / n);
		}
		for (int i = 0; i < p->m; ++i) {
			uint32_t r = 0;
			for (int b = 0; b < p->logm; ++b) {
				if (i & (1 << b)) {
					r |= 1u << (p->logm - 1 - b);
				}
			}
			p->rev[i] = r;
		}
	} else {
		p->sp_re = (real_t*)malloc (sizeof (real_t) * n);
		p->sp_im = (real_t*)malloc (sizeof (real_t) * n);
		for (int k = 0; k < n; ++k) {
			p->sp_re[k] = (real_t)cos (twopi * k / n);
			p->sp_im[k] = (real_t)-sin (twopi * k / n);
		}
	}
	return p;
}

fftwf_plan
fftwf_plan_dft_r2c_1d (int n, float* in, fftwf_complex* out, unsigned flags)
{
	(void)in;
	(void)out;
	(void)flags;
	return plan_new (n, 0);
}

fftwf_plan
fftwf_plan_dft_c2r_1d (int n, fftwf_complex* in, float* out, unsigned flags)
{
	(void)in;
	(void)out;
	(void)flags;
	return plan_new (n, 1);
}

void
fftwf_destroy_plan (fftwf_plan p)
{
	if (!p) {
		return;
	}
	free (p->tw_re);
	free (p->tw_im);
	free (p->sp_re);
	free (p->sp_im);
	free (p->rev);
	free (p->wr);
	free (p->wi);
	free (p);
}

void
fftwf_cleanup (void)
{
}

/* in-place complex FFT of size m on (wr, wi), data already bit-reversed.
 * sign = -1 forward, +1 backward (conjugated twiddles). */
static void
cfft_core (const struct standin_fftwf_plan_s* p, real_t* wr, real_t* wi, int sign)
{
	const int m = p->m;
	for (int half = 1; half < m; half <<= 1) {
		const int step = m / (2 * half);
		for (int base = 0; base < m; base += 2 * half) {
			for (int j = 0; j < half; ++j) {
				const real_t c  = p->tw_re[j * step];
				const real_t s  = sign < 0 ? p->tw_im[j * step] : -p->tw_im[j * step];
				const int    a  = base + j;
				const int    b  = a + half;
				const real_t tr = wr[b] * c - wi[b] * s;
				const real_t ti = wr[b] * s + wi[b] * c;
				wr[b]           = wr[a] - tr;
				wi[b]           = wi[a] - ti;
				wr[a] += tr;
				wi[a] += ti;
			}
		}
	}
}

void
fftwf_execute_dft_r2c (const fftwf_plan p, float* in, fftwf_complex* out)
{
	const int n = p->n;
	if (!p->pow2) {
		for (int k = 0; k <= n / 2; ++k) {
			real_t sr = 0, si = 0;
			for (int j = 0; j < n; ++j) {
				const int t = (int)(((int64_t)j * k) % n);
				sr += (real_t)in[j] * p->sp_re[t];
				si += (real_t)in[j] * p->sp_im[t];
			}
			out[k][0] = (float)sr;
			out[k][1] = (float)si;
		}
		return;
	}
	const int m  = p->m;
	real_t*   wr = p->wr;
	real_t*   wi = p->wi;
	for (int i = 0; i < m; ++i) {
		const uint32_t r = p->rev[i];
		wr[r]            = (real_t)in[2 * i];
		wi[r]            = (real_t)in[2 * i + 1];
	}
	cfft_core (p, wr, wi, -1);
	/* split: X[k] = E[k] + W^k O[k],  E = (Z[k] + conj Z[m-k])/2, O = (Z[k] - conj Z[m-k])/(2i) */
	for (int k = 0; k <= m; ++k) {
		const int    ka = k % m;
		const int    kb = (m - k) % m;
		const real_t er = (real_t)0.5 * (wr[ka] + wr[kb]);
		const real_t ei = (real_t)0.5 * (wi[ka] - wi[kb]);
		const real_t or_ = (real_t)0.5 * (wi[ka] + wi[kb]);
		const real_t oi = (real_t)-0.5 * (wr[ka] - wr[kb]);
		const real_t c  = p->sp_re[k];
		const real_t s  = p->sp_im[k];
		out[k][0]       = (float)(er + or_ * c - oi * s);
		out[k][1]       = (float)(ei + or_ * s + oi * c);
	}
}

void
fftwf_execute_dft_c2r (const fftwf_plan p, fftwf_complex* in, float* out)
{
	const int n = p->n;
	if (!p->pow2) {
		const int h = n / 2;
		for (int j = 0; j < n; ++j) {
			real_t acc = (real_t)in[0][0];
			for (int k = 1; k <= h; ++k) {
				const int t = (int)(((int64_t)j * k) % n);
				/* exp(+i phi) = (sp_re, -sp_im) */
				const real_t c = p->sp_re[t];
				const real_t s = -p->sp_im[t];
				if ((n % 2 == 0) && k == h) {
					acc += (real_t)in[k][0] * c;
				} else {
					acc += (real_t)2 * ((real_t)in[k][0] * c - (real_t)in[k][1] * s);
				}
			}
			out[j] = (float)acc;
		}
		return;
	}
	const int m  = p->m;
	real_t*   wr = p->wr;
	real_t*   wi = p->wi;
	for (int k = 0; k < m; ++k) {
		/* Hermitian input: imaginary parts of DC and Nyquist are ignored */
		const real_t ar = (real_t)in[k][0];
		const real_t ai = (k == 0) ? (real_t)0 : (real_t)in[k][1];
		const real_t br = (real_t)in[m - k][0];
		const real_t bi = (k == 0) ? (real_t)0 : -(real_t)in[m - k][1]; /* conj X[m-k]; k==0 -> Nyquist */
		const real_t sr = ar + br;
		const real_t si = ai + bi;
		const real_t dr = ar - br;
		const real_t di = ai - bi;
		/* i * exp(+2 pi i k / n) * d */
		const real_t c  = p->sp_re[k];
		const real_t s  = -p->sp_im[k];
		const real_t tr = dr * c - di * s;
		const real_t ti = dr * s + di * c;
		const uint32_t r = p->rev[k];
		wr[r]            = sr - ti;
		wi[r]            = si + tr;
	}
	cfft_core (p, wr, wi, +1);
	for (int j = 0; j < m; ++j) {
		out[2 * j]     = (float)wr[j];
		out[2 * j + 1] = (float)wi[j];
	}
}
