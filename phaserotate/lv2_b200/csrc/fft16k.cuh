// 16384-point complex FFT convolution core of fftconv_kernel (sm_100a).
//
// One CTA of 512 threads transforms one segment of M = 16384 complex points
// that lives in shared memory.  M = 32 * 32 * 16; a point index is
//     i = i2 * 512 + i1 * 16 + i0,   i2, i1 < 32, i0 < 16.
// Forward (decimation in frequency), every butterfly entirely in registers:
//   P1   radix-32 over i2 (stride 512), one butterfly per thread e = (i1, i0),
//        input straight from global memory, output q1 times W_M^(e q1)
//   P2   radix-32 over i1 (stride 16), one butterfly per thread (q1, i0)
//   MID  per row (q1, q2) of 16 contiguous points: times W_512^(i0 q2), radix-16
//        over i0 -> q3, times the filter spectrum G[q1 + 32 q2 + 1024 q3], and
//        straight back: inverse radix-16, times conj W_512^(i0 q2)
// Inverse: P2 again with the conjugate roots, then P1' = conj twiddle + inverse
// radix-32 over q1, which leaves the outputs i = e + 512 k, k < 32, in the
// registers of thread e for the epilogue.  Four shared-memory round trips per
// segment (the L1/shared data pipe and the FMA pipe are the two co-critical
// resources).
//
// Shared-memory layout: row = i >> 4 (16 points = 128 bytes), 16-byte chunk
// index XOR-swizzled with ((row >> 5) ^ row) & 7.  Every access pattern of the
// five passes is then bank-conflict free without padding: half-warps touch one
// whole row (P1, P2), quarter-warps touch one chunk of eight consecutive rows
// whose swizzle keys differ (MID, 128-bit accesses).
//
// All complex arithmetic uses the packed fp32 instructions of sm_100 (FADD2 /
// FMUL2 / FFMA2 on an aligned register pair = one complex number); component
// swaps and negations compile to operand modifiers (.LO_HI, .NP, -R, .F32
// broadcast), not instructions.  The same source compiles for the host with
// scalar arithmetic so that the index algebra is unit-tested without a GPU
// (tools/fft16k_hosttest.cu, tests/test_fft16k_host.py).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PRK_HD __host__ __device__ __forceinline__
#else
#define PRK_HD inline
#endif

namespace prk {

constexpr int kLog2M       = 14;
constexpr int kM           = 1 << kLog2M; // complex FFT size per segment
constexpr int kConvThreads = 512;
constexpr int kTwP1Rows    = 10;          // W_M^(e ql), ql = 1..3 | W_M^(4 e qh), qh = 1..7
constexpr int kTwMidRows   = 15;          // W_512^(j q2), j = 1..15

// ---------------------------------------------------------------------------
// complex helpers
// ---------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
PRK_HD float2 cadd (float2 a, float2 b) { return __fadd2_rn (a, b); }
PRK_HD float2 csub (float2 a, float2 b) { return __fadd2_rn (a, make_float2 (-b.x, -b.y)); }
PRK_HD float2 caddi (float2 a, float2 b) { return __fadd2_rn (a, make_float2 (-b.y, b.x)); } // a + i b
PRK_HD float2 csubi (float2 a, float2 b) { return __fadd2_rn (a, make_float2 (b.y, -b.x)); } // a - i b
// a * b = b.x * a + b.y * (i a): the factor b enters only through broadcasts of its
// two components and the data a through a swap / negate modifier, so neither
// product needs a rearranged copy of b (and a * conj(b) needs no negated one).
// Operand order matters: the half-negating modifier (.NP) exists only for the
// first source operand, the .F32 broadcast for either, and the compiler does not
// commute the operands (probed with cuobjdump; a swapped order costs an extra
// FADD + MOV per product).
PRK_HD float2 cmul (float2 a, float2 b)
{
	return __ffma2_rn (make_float2 (-a.y, a.x), make_float2 (b.y, b.y), __fmul2_rn (a, make_float2 (b.x, b.x)));
}
PRK_HD float2 cmulc (float2 a, float2 b) // a * conj(b) = b.x * a - b.y * (i a)
{
	return __ffma2_rn (make_float2 (a.y, -a.x), make_float2 (b.y, b.y), __fmul2_rn (a, make_float2 (b.x, b.x)));
}
#else
PRK_HD float2 cadd (float2 a, float2 b) { return make_float2 (a.x + b.x, a.y + b.y); }
PRK_HD float2 csub (float2 a, float2 b) { return make_float2 (a.x - b.x, a.y - b.y); }
PRK_HD float2 caddi (float2 a, float2 b) { return make_float2 (a.x - b.y, a.y + b.x); }
PRK_HD float2 csubi (float2 a, float2 b) { return make_float2 (a.x + b.y, a.y - b.x); }
PRK_HD float2 cmul (float2 a, float2 b) { return make_float2 (a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
PRK_HD float2 cmulc (float2 a, float2 b) { return make_float2 (a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }
#endif
// a * (wr + i wi), (wr, wi) compile-time constants: written as wr * a + wi * (i a) so
// that both constants are 32-bit immediates broadcast to the two halves of a
// packed instruction (no registers, no constant materialisation)
#if defined(__CUDA_ARCH__)
PRK_HD float2 cmulk (float2 a, float wr, float wi)
{
	return __ffma2_rn (make_float2 (-a.y, a.x), make_float2 (wi, wi), __fmul2_rn (a, make_float2 (wr, wr)));
}
#else
PRK_HD float2 cmulk (float2 a, float wr, float wi) { return make_float2 (a.x * wr - a.y * wi, a.y * wr + a.x * wi); }
#endif

// s * a + c and s * (i a) + c, s a scalar (compile-time constant after unrolling:
// a 32-bit immediate broadcast to both halves of one FFMA2)
#if defined(__CUDA_ARCH__)
PRK_HD float2 cfma (float2 a, float s, float2 c) { return __ffma2_rn (a, make_float2 (s, s), c); }
PRK_HD float2 cfmai (float2 a, float s, float2 c) { return __ffma2_rn (make_float2 (-a.y, a.x), make_float2 (s, s), c); }
#else
PRK_HD float2 cfma (float2 a, float s, float2 c) { return make_float2 (a.x * s + c.x, a.y * s + c.y); }
PRK_HD float2 cfmai (float2 a, float s, float2 c) { return make_float2 (-a.y * s + c.x, a.x * s + c.y); }
#endif

// DIR = -1 forward (roots exp(-2 pi i / n)), +1 inverse
template <int DIR>
PRK_HD void dft4 (float2& a0, float2& a1, float2& a2, float2& a3)
{
	const float2 s02 = cadd (a0, a2), d02 = csub (a0, a2);
	const float2 s13 = cadd (a1, a3), d13 = csub (a1, a3);
	a0 = cadd (s02, s13);
	a2 = csub (s02, s13);
	a1 = DIR < 0 ? csubi (d02, d13) : caddi (d02, d13);
	a3 = DIR < 0 ? caddi (d02, d13) : csubi (d02, d13);
}

// cos / sin (2 pi m / 32), m = 0..31
PRK_HD constexpr float cos32 (int m)
{
	constexpr float c[9] = { 1.f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f, 0.70710678118654752440f,
		                     0.55557023301960222474f, 0.38268343236508977173f, 0.19509032201612826785f, 0.f };
	m &= 31;
	if (m > 16) m = 32 - m;
	return m <= 8 ? c[m] : -c[16 - m];
}
PRK_HD constexpr float sin32 (int m) { return cos32 (m - 8); }

// a * exp(DIR 2 pi i m / 32)
template <int DIR, int m>
PRK_HD float2 mulw32 (float2 a)
{
	constexpr int mm = m & 31;
	if (mm == 0) return a;
	if (mm == 8) return DIR < 0 ? make_float2 (a.y, -a.x) : make_float2 (-a.y, a.x);
	if (mm == 16) return make_float2 (-a.x, -a.y);
	if (mm == 24) return DIR < 0 ? make_float2 (-a.y, a.x) : make_float2 (a.y, -a.x);
	return cmulk (a, cos32 (mm), (DIR < 0 ? -1.f : 1.f) * sin32 (mm));
}

// Twiddled radix-2 butterfly  p = x + W b,  q = x - W b,  W = exp(DIR 2 pi i m / 32),
// m a loop constant after unrolling.  A twiddle that is consumed only by such a
// +/- pair never needs its own complex product (FMUL2 + FFMA2): with
// W = wr (1 + i t), t = wi / wr,
//     c = b + t (i b)          one FFMA2
//     p = x + wr c, q = x - wr c   two FFMA2
// i.e. three packed instructions where product + two adds take four; when
// |wi| > |wr| the same with W = i wi (1 + i t), t = -wr / wi, so |t| <= 1 always.
template <int DIR>
PRK_HD void bfly_w (float2 x, float2 b, int m, float2& p, float2& q)
{
	m &= 31;
	if (m == 0) {
		p = cadd (x, b);
		q = csub (x, b);
	} else if (m == 16) {
		p = csub (x, b);
		q = cadd (x, b);
	} else if (m == 8 || m == 24) {
		const bool plus_i = (m == 8) == (DIR > 0); // W = +i
		p = plus_i ? caddi (x, b) : csubi (x, b);
		q = plus_i ? csubi (x, b) : caddi (x, b);
	} else {
		const float wr = cos32 (m), wi = (DIR < 0 ? -1.f : 1.f) * sin32 (m);
		if ((wr < 0 ? -wr : wr) >= (wi < 0 ? -wi : wi)) {
			const float2 c = cfmai (b, wi / wr, b);
			p              = cfma (c, wr, x);
			q              = cfma (c, -wr, x);
		} else {
			const float2 c = cfmai (b, -wr / wi, b);
			p              = cfmai (c, wi, x);
			q              = cfmai (c, -wi, x);
		}
	}
}

// dft4 whose inputs a2, a3 still lack their twiddles exp(DIR 2 pi i m / 32), m = m2, m3
template <int DIR>
PRK_HD void dft4_lazy (float2& a0, float2& a1, float2& a2, float2& a3, int m2, int m3)
{
	float2 s02, d02, s13, d13;
	bfly_w<DIR> (a0, a2, m2, s02, d02);
	bfly_w<DIR> (a1, a3, m3, s13, d13);
	a0 = cadd (s02, s13);
	a2 = csub (s02, s13);
	a1 = DIR < 0 ? csubi (d02, d13) : caddi (d02, d13);
	a3 = DIR < 0 ? caddi (d02, d13) : csubi (d02, d13);
}

// 8-point DFT, natural order in and out; inputs v[4 .. 7] still lack the
// twiddles exp(DIR 2 pi i m / 32), m = m4 .. m7 (0 = none)
template <int DIR>
PRK_HD void dft8 (float2 (&v)[8], int m4 = 0, int m5 = 0, int m6 = 0, int m7 = 0)
{
	float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
	float2 o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
	dft4_lazy<DIR> (e0, e1, e2, e3, m4, m6);
	dft4_lazy<DIR> (o0, o1, o2, o3, m5, m7);
	v[0] = cadd (e0, o0);
	v[4] = csub (e0, o0);
	bfly_w<DIR> (e1, o1, 4, v[1], v[5]);
	bfly_w<DIR> (e2, o2, 8, v[2], v[6]);
	bfly_w<DIR> (e3, o3, 12, v[3], v[7]);
}

// 16-point DFT, natural order in and out (4 x 4)
template <int DIR>
PRK_HD void dft16 (float2 (&u)[16])
{
	float2 t[4][4]; // t[k0][ql]
#pragma unroll
	for (int k0 = 0; k0 < 4; ++k0) {
		float2 a0 = u[k0], a1 = u[k0 + 4], a2 = u[k0 + 8], a3 = u[k0 + 12];
		dft4<DIR> (a0, a1, a2, a3);
		t[k0][0] = a0;
		t[k0][1] = a1;
		t[k0][2] = a2;
		t[k0][3] = a3;
	}
	// twiddles W_16^(k0 ql) = W_32^(2 k0 ql): k0 = 1 applied here, k0 = 2, 3 inside the
	// butterflies that consume them (bfly_w)
	t[1][1] = mulw32<DIR, 2> (t[1][1]);
	t[1][2] = mulw32<DIR, 4> (t[1][2]);
	t[1][3] = mulw32<DIR, 6> (t[1][3]);
#pragma unroll
	for (int ql = 0; ql < 4; ++ql) {
		float2 a0 = t[0][ql], a1 = t[1][ql], a2 = t[2][ql], a3 = t[3][ql];
		dft4_lazy<DIR> (a0, a1, a2, a3, 4 * ql, 6 * ql); // a[qh] = y[ql + 4 qh]
		u[ql]      = a0;
		u[ql + 4]  = a1;
		u[ql + 8]  = a2;
		u[ql + 12] = a3;
	}
}

// a * exp(DIR 2 pi i m / 32); m is a loop constant after unrolling
template <int DIR>
PRK_HD float2 mulw32v (float2 a, int m)
{
	m &= 31;
	if (m == 0) return a;
	if (m == 8) return DIR < 0 ? make_float2 (a.y, -a.x) : make_float2 (-a.y, a.x);
	if (m == 16) return make_float2 (-a.x, -a.y);
	if (m == 24) return DIR < 0 ? make_float2 (-a.y, a.x) : make_float2 (a.y, -a.x);
	return cmulk (a, cos32 (m), (DIR < 0 ? -1.f : 1.f) * sin32 (m));
}

// 32-point DFT in two register stages, written so that a pass can stream:
//   "4 x 8" (input n = k0 + 8 k1, output q = ql + 4 qh)
//      stage 1, per k0: radix-4 over k1, constant twiddle W_32^(k0 ql)      -> t[k0][ql]
//      stage 2, per ql: radix-8 over k0                                      -> y[ql + 4 qh]
//   "8 x 4" (input n = n0 + 4 n1, output p = p0 + 8 p1)
//      stage 1, per n0: radix-8 over n1, constant twiddle W_32^(n0 p0)      -> t[p0][n0]
//      stage 2, per p0: radix-4 over n0                                      -> y[p0 + 8 p1]
// A pass issues its loads in the order
// stage 1 consumes them and stores each stage-2 group as soon as it exists, so
// that the butterflies run while the rest of the data is still in flight.
// The constant twiddles between the two stages are applied in stage 1 only where
// stage 2 consumes the value as the first operand of a butterfly (k0 < 4, n0 < 2);
// the others ride on the stage-2 butterflies (bfly_w): 196 packed instructions.
template <int DIR>
PRK_HD void bf48_stage1 (float2 (&t)[8][4], int k0, float2 a0, float2 a1, float2 a2, float2 a3)
{
	dft4<DIR> (a0, a1, a2, a3);
	t[k0][0] = a0;
	t[k0][1] = k0 < 4 ? mulw32v<DIR> (a1, k0) : a1;
	t[k0][2] = k0 < 4 ? mulw32v<DIR> (a2, 2 * k0) : a2;
	t[k0][3] = k0 < 4 ? mulw32v<DIR> (a3, 3 * k0) : a3;
}
template <int DIR>
PRK_HD void bf48_stage2 (const float2 (&t)[8][4], int ql, float2 (&v)[8])
{
#pragma unroll
	for (int k0 = 0; k0 < 8; ++k0) v[k0] = t[k0][ql];
	dft8<DIR> (v, 4 * ql, 5 * ql, 6 * ql, 7 * ql); // v[qh] = y[ql + 4 qh]
}
template <int DIR>
PRK_HD void bf84_stage1 (float2 (&t)[8][4], int n0, float2 (&v)[8])
{
	dft8<DIR> (v); // v[p0]
#pragma unroll
	for (int p0 = 0; p0 < 8; ++p0) t[p0][n0] = n0 < 2 ? mulw32v<DIR> (v[p0], n0 * p0) : v[p0];
}
template <int DIR>
PRK_HD void bf84_stage2 (const float2 (&t)[8][4], int p0, float2& y0, float2& y1, float2& y2, float2& y3)
{
	y0 = t[p0][0];
	y1 = t[p0][1];
	y2 = t[p0][2];
	y3 = t[p0][3];
	dft4_lazy<DIR> (y0, y1, y2, y3, 2 * p0, 3 * p0); // y[p1] = out[p0 + 8 p1]
}

// natural order in and out (used by the microbenchmarks and host tests)
template <int DIR>
PRK_HD void dft32 (float2 (&u)[32])
{
	float2 t[8][4];
#pragma unroll
	for (int k0 = 0; k0 < 8; ++k0) bf48_stage1<DIR> (t, k0, u[k0], u[k0 + 8], u[k0 + 16], u[k0 + 24]);
#pragma unroll
	for (int ql = 0; ql < 4; ++ql) {
		float2 v[8];
		bf48_stage2<DIR> (t, ql, v);
#pragma unroll
		for (int qh = 0; qh < 8; ++qh) u[ql + 4 * qh] = v[qh];
	}
}

// ---------------------------------------------------------------------------
// shared-memory layout
// ---------------------------------------------------------------------------
// float2 index of point (row, col), row = i >> 4, col = i & 15
PRK_HD int swz (int row, int col) { return (row << 4) | (((((col >> 1) ^ (row >> 5) ^ row) & 7) << 1) | (col & 1)); }

// Points (row0 + 32 q, col), q = 0..31 (P1: row0 = e >> 4 < 32) or (row0 + k, col),
// k = 0..31 (P2: row0 = 32 q1) differ from one another only by a multiple of the
// stride plus a swizzle term that depends on (q or k) & 7: eight offsets per
// thread, every access an immediate offset from one of them.
struct SwzCol {
	int x[8];
};
PRK_HD SwzCol swz_col (int row0, int col)
{
	SwzCol s;
	const int a = ((col >> 1) ^ (row0 >> 5) ^ row0) & 7;
#pragma unroll
	for (int c = 0; c < 8; ++c) s.x[c] = (row0 << 4) | ((a ^ c) << 1) | (col & 1);
	return s;
}

#if defined(__CUDA_ARCH__)
#define PRK_LDG(p) __ldg (p)
#else
#define PRK_LDG(p) (*(p))
#endif

// The 31 twiddles W_M^(e q), q = ql + 4 qh, from ten table entries:
// rows 0..2 = W_M^(e ql), ql = 1..3; rows 3..9 = W_M^(4 e qh), qh = 1..7.
struct TwP1 {
	float2 tl[4], th[8]; // index 0 unused
};
PRK_HD TwP1 load_tw_p1 (const float2* __restrict__ tw, int e)
{
	TwP1 t;
	t.tl[0] = t.th[0] = make_float2 (1.f, 0.f);
#pragma unroll
	for (int a = 1; a < 4; ++a) t.tl[a] = PRK_LDG (tw + (a - 1) * 512 + e);
#pragma unroll
	for (int b = 1; b < 8; ++b) t.th[b] = PRK_LDG (tw + (2 + b) * 512 + e);
	return t;
}
template <bool CONJ>
PRK_HD float2 apply_tw_p1 (float2 v, const TwP1& t, int ql, int qh)
{
	if (ql && qh) {
		const float2 w = cmul (t.tl[ql], t.th[qh]);
		return CONJ ? cmulc (v, w) : cmul (v, w);
	}
	if (ql) return CONJ ? cmulc (v, t.tl[ql]) : cmul (v, t.tl[ql]);
	if (qh) return CONJ ? cmulc (v, t.th[qh]) : cmul (v, t.th[qh]);
	return v;
}

// P1: global -> registers -> radix-32 over i2 -> twiddle -> shared.  Thread e.
// The loads are issued in two batches of 16 (k0 < 4, k0 >= 4): a stereo frame
// pair arrives as 16 bytes of which 8 are used, so 32 loads in flight would
// need the whole register file.
struct NoStash {
	PRK_HD void operator() (int, float2, float2, float2, float2) const {}
};
// `reuse (k, v)`: if the raw inputs k .. k + 3 (k a multiple of 4) are already at
// hand (overlap with the previous segment of the same stream), deliver them and
// return true
struct NoReuse {
	PRK_HD bool operator() (int, float2 (&)[4]) const { return false; }
};
// `stash (k, a, b, c, d)` receives the raw inputs k .. k + 3 (k a multiple of 4)
template <class Loader, class Stash = NoStash, class Reuse = NoReuse>
PRK_HD void p1_forward (float2* sm, const float2* __restrict__ tw, int e, const Loader& ld, const Stash& stash = Stash (), const Reuse& reuse = Reuse ())
{
	float2     t[8][4];
	const auto lt = ld.thread (e);
#pragma unroll
	for (int half = 0; half < 2; ++half) {
		float2 u[4][4];
#pragma unroll
		for (int k1 = 0; k1 < 4; ++k1) {
			float2 v[4];
			if (reuse (4 * half + 8 * k1, v)) {
#pragma unroll
				for (int kk = 0; kk < 4; ++kk) u[kk][k1] = v[kk];
			} else {
#pragma unroll
				for (int kk = 0; kk < 4; ++kk) u[kk][k1] = lt (512 * (4 * half + kk + 8 * k1));
			}
		}
#pragma unroll
		for (int k1 = 0; k1 < 4; ++k1) stash (4 * half + 8 * k1, u[0][k1], u[1][k1], u[2][k1], u[3][k1]);
#pragma unroll
		for (int kk = 0; kk < 4; ++kk) bf48_stage1<-1> (t, 4 * half + kk, u[kk][0], u[kk][1], u[kk][2], u[kk][3]);
	}
	const TwP1   w  = load_tw_p1 (tw, e);
	const SwzCol so = swz_col (e >> 4, e & 15); // row = 32 q + (e >> 4): key term q & 7
#pragma unroll
	for (int ql = 0; ql < 4; ++ql) {
		float2 v[8];
		bf48_stage2<-1> (t, ql, v);
#pragma unroll
		for (int qh = 0; qh < 8; ++qh) {
			const int q                  = ql + 4 * qh;
			sm[q * 512 + so.x[q & 7]] = apply_tw_p1<false> (v[qh], w, ql, qh);
		}
	}
}

// P2: radix-32 over i1 (forward) / q2 (inverse) for fixed (q1, i0).  Warp wp
// handles the two 512-point blocks q1 = wp (lanes 0..15, i0 = lane) and
// q1 = wp + 16 (lanes 16..31) - the same two blocks whose rows it owns in MID
// (rows t and t + 512), so P2 -> MID -> P2' of a block pair runs inside one warp
// with warp-level synchronisation only and the warps drift apart: while one
// waits for shared memory another one has the FMA pipe.
PRK_HD int p2_block (int t) { return (t >> 5) + 16 * ((t >> 4) & 1); }
template <int DIR>
PRK_HD void p2_pass (float2* sm, int t)
{
	const SwzCol so = swz_col (p2_block (t) * 32, t & 15); // row = 32 q1 + k: key term k & 7
	float2       s1[8][4];
	{
		float2 u[8][4];
#pragma unroll
		for (int k0 = 0; k0 < 8; ++k0) {
#pragma unroll
			for (int k1 = 0; k1 < 4; ++k1) u[k0][k1] = sm[(k0 + 8 * k1) * 16 + so.x[k0]];
		}
#pragma unroll
		for (int k0 = 0; k0 < 8; ++k0) bf48_stage1<DIR> (s1, k0, u[k0][0], u[k0][1], u[k0][2], u[k0][3]);
	}
#pragma unroll
	for (int ql = 0; ql < 4; ++ql) {
		float2 v[8];
		bf48_stage2<DIR> (s1, ql, v);
#pragma unroll
		for (int qh = 0; qh < 8; ++qh) sm[(ql + 4 * qh) * 16 + so.x[(ql + 4 * qh) & 7]] = v[qh];
	}
}

// MID: rows t and t + 512 (same q2 = t & 31): twiddle, radix-16, filter, inverse
// radix-16, conjugate twiddle.  The filter spectrum of a row, already scaled by
// 1 / M, comes from `gs`:  gs.issue (h) starts fetching the 16 bins
// G[f(row, q3)], f(row, q3) = (row >> 5) + 32 (row & 31) + 1024 q3, of row
// t + 512 h and gs.get (g) delivers them as g[c] = (G[f(row, 2c)], G[f(row, 2c + 1)]).
// On the device that is the thread's Tensor Memory stash (kernels.cuh), in the
// host emulation and the microbenchmarks the table itself.
struct GTable { // G4[c * 1024 + row] as built by make_filter_spectrum()
	const float4* G4;
	int           t, row;
	PRK_HD void issue (int h) { row = t + 512 * h; }
	PRK_HD void get (float4 (&g)[8]) const
	{
#pragma unroll
		for (int c = 0; c < 8; ++c) g[c] = PRK_LDG (G4 + c * 1024 + row);
	}
};
// Two-partition form (FIR longer than M/2 half-taps, i.e. CLI block 32768): the
// taps are split in two halves of Lp = M/2, the segment advance V equals Lp, so
// partition 1 of output segment s is partition-1-filter times the spectrum of
// input segment s - 1:  out_s = IFFT (Z_s G0 + Z_{s-1} G1).  A thread owns the
// same two spectrum rows in every segment, so Z_{s-1} is thread-private scratch
// (`scr`, global memory, L2 resident: element (h * 8 + c) * 512 of the thread's
// base holds chunk c of row t + 512 h).
//   MID_CONV      one partition (the normal case)
//   MID_SPECTRUM  forward half only: leave Z in `scr` (first segment of a run)
//   MID_CONV2     read Z_{s-1} from `scr`, replace it by Z_s, two-partition product
enum { MID_CONV = 0, MID_SPECTRUM = 1, MID_CONV2 = 2 };
#if defined(__CUDA_ARCH__)
#define PRK_LDCG(p) __ldcg (p)
#define PRK_STCG(p, v) __stcg (p, v)
#else
#define PRK_LDCG(p) (*(p))
#define PRK_STCG(p, v) (*(p) = (v))
#endif
template <int MODE = MID_CONV, class GSrc>
PRK_HD void mid_pass (float2* sm, GSrc gs, const float2* __restrict__ twm, int t, float4* scr = nullptr, const float4* __restrict__ G0 = nullptr,
                      const float4* __restrict__ G1 = nullptr)
{
	const int q2 = t & 31;
	float2    tw[16];
#pragma unroll
	for (int j = 1; j < 16; ++j) tw[j] = PRK_LDG (twm + (j - 1) * 32 + q2);
#pragma unroll
	for (int h = 0; h < 2; ++h) {
		const int row = t + 512 * h;
		const int s   = ((row >> 5) ^ row) & 7;
		float4*   rp  = reinterpret_cast<float4*> (sm + (row << 4));
		if (MODE == MID_CONV) gs.issue (h);
		float2 u[16];
#pragma unroll
		for (int c = 0; c < 8; ++c) {
			const float4 v = rp[c ^ s];
			u[2 * c]       = make_float2 (v.x, v.y);
			u[2 * c + 1]   = make_float2 (v.z, v.w);
		}
#pragma unroll
		for (int j = 1; j < 16; ++j) u[j] = cmul (u[j], tw[j]);
		dft16<-1> (u);
		if (MODE == MID_SPECTRUM) {
#pragma unroll
			for (int c = 0; c < 8; ++c) PRK_STCG (scr + (h * 8 + c) * 512, make_float4 (u[2 * c].x, u[2 * c].y, u[2 * c + 1].x, u[2 * c + 1].y));
			continue;
		}
		if (MODE == MID_CONV2) {
#pragma unroll
			for (int c = 0; c < 8; ++c) {
				const float4 pv = PRK_LDCG (scr + (h * 8 + c) * 512);
				const float4 g0 = PRK_LDG (G0 + c * 1024 + row), g1 = PRK_LDG (G1 + c * 1024 + row);
				PRK_STCG (scr + (h * 8 + c) * 512, make_float4 (u[2 * c].x, u[2 * c].y, u[2 * c + 1].x, u[2 * c + 1].y));
				u[2 * c]     = cadd (cmul (u[2 * c], make_float2 (g0.x, g0.y)), cmul (make_float2 (pv.x, pv.y), make_float2 (g1.x, g1.y)));
				u[2 * c + 1] = cadd (cmul (u[2 * c + 1], make_float2 (g0.z, g0.w)), cmul (make_float2 (pv.z, pv.w), make_float2 (g1.z, g1.w)));
			}
		} else {
			float4 g[8];
			gs.get (g);
#pragma unroll
			for (int c = 0; c < 8; ++c) {
				u[2 * c]     = cmul (u[2 * c], make_float2 (g[c].x, g[c].y));
				u[2 * c + 1] = cmul (u[2 * c + 1], make_float2 (g[c].z, g[c].w));
			}
		}
		dft16<+1> (u);
#pragma unroll
		for (int j = 1; j < 16; ++j) u[j] = cmulc (u[j], tw[j]);
		// two 8-byte stores per chunk: a 16-byte store would need both results in
		// one aligned register quad, which costs moves
		float2* rp2 = reinterpret_cast<float2*> (rp);
#pragma unroll
		for (int c = 0; c < 8; ++c) {
			rp2[2 * (c ^ s)]     = u[2 * c];
			rp2[2 * (c ^ s) + 1] = u[2 * c + 1];
		}
	}
}

// P1': shared -> conjugate twiddle -> inverse radix-32 over q1.  On return w[k]
// is output point i = e + 512 k of the segment.
PRK_HD void p1_inverse (const float2* sm, const float2* __restrict__ tw, int e, float2 (&w)[32])
{
	const TwP1   tf = load_tw_p1 (tw, e);
	const SwzCol so = swz_col (e >> 4, e & 15);
	float2       s1[8][4];
	{
		float2 u[4][8];
#pragma unroll
		for (int n0 = 0; n0 < 4; ++n0) {
#pragma unroll
			for (int n1 = 0; n1 < 8; ++n1) u[n0][n1] = sm[(n0 + 4 * n1) * 512 + so.x[(n0 + 4 * n1) & 7]];
		}
#pragma unroll
		for (int n0 = 0; n0 < 4; ++n0) {
			float2 v[8];
#pragma unroll
			for (int n1 = 0; n1 < 8; ++n1) v[n1] = apply_tw_p1<true> (u[n0][n1], tf, n0, n1);
			bf84_stage1<+1> (s1, n0, v);
		}
	}
#pragma unroll
	for (int p0 = 0; p0 < 8; ++p0) bf84_stage2<+1> (s1, p0, w[p0], w[p0 + 8], w[p0 + 16], w[p0 + 24]);
}

} // namespace prk
