// Host-side construction of the constant tables of the 16384-point FFT
// convolution (fft16k.cuh): filter spectrum in MID-pass order and twiddles.
#pragma once

#include <cmath>
#include <vector>

#include "fft16k.cuh"

namespace prk {

// in-place radix-2 complex FFT (double), n a power of two; sign -1 forward
inline void
host_fft (std::vector<double>& re, std::vector<double>& im, int sign)
{
	const size_t n = re.size ();
	for (size_t i = 1, j = 0; i < n; ++i) {
		size_t bit = n >> 1;
		for (; j & bit; bit >>= 1) {
			j ^= bit;
		}
		j ^= bit;
		if (i < j) {
			std::swap (re[i], re[j]);
			std::swap (im[i], im[j]);
		}
	}
	for (size_t len = 2; len <= n; len <<= 1) {
		const double ang = sign * 2.0 * M_PI / (double)len;
		for (size_t i = 0; i < n; i += len) {
			for (size_t k = 0; k < len / 2; ++k) {
				const double wr = std::cos (ang * (double)k), wi = std::sin (ang * (double)k);
				const size_t a = i + k, b = i + k + len / 2;
				const double tr = re[b] * wr - im[b] * wi, ti = re[b] * wi + im[b] * wr;
				re[b] = re[a] - tr;
				im[b] = im[a] - ti;
				re[a] += tr;
				im[a] += ti;
			}
		}
	}
}

// Spectrum of the real taps g[0..n_taps) (zero-padded to kM), divided by kM,
// in the order mid_pass() reads it: G4[c * 1024 + row] = (G[f(row, 2c)], G[f(row, 2c+1)]),
// f(row, q3) = (row >> 5) + 32 (row & 31) + 1024 q3.  Returned as kM float2.
inline std::vector<float2>
make_filter_spectrum (const float* g, int n_taps)
{
	std::vector<double> re ((size_t)kM, 0.0), im ((size_t)kM, 0.0);
	for (int j = 0; j < n_taps; ++j) {
		re[(size_t)j] = (double)g[j];
	}
	host_fft (re, im, -1);
	std::vector<float2> G ((size_t)kM);
	for (int c = 0; c < 8; ++c) {
		for (int row = 0; row < 1024; ++row) {
			for (int h = 0; h < 2; ++h) {
				const int q3 = 2 * c + h;
				const int f  = (row >> 5) + 32 * (row & 31) + 1024 * q3;
				G[(size_t)(2 * (c * 1024 + row) + h)] = make_float2 ((float)(re[(size_t)f] / kM), (float)(im[(size_t)f] / kM));
			}
		}
	}
	return G;
}

// [kTwP1Rows][512] : rows 0..2 = W_M^(e ql), ql = 1..3; rows 3..9 = W_M^(4 e qh), qh = 1..7
// followed by [kTwMidRows][32] : W_512^(j q2), j = 1..15
inline std::vector<float2>
make_twiddles ()
{
	std::vector<float2> tw;
	for (int r = 0; r < kTwP1Rows; ++r) {
		const int mult = r < 3 ? r + 1 : 4 * (r - 2);
		for (int e = 0; e < 512; ++e) {
			const double a = -2.0 * M_PI * (double)((long long)e * mult) / (double)kM;
			tw.push_back (make_float2 ((float)std::cos (a), (float)std::sin (a)));
		}
	}
	for (int j = 1; j < 16; ++j) {
		for (int q2 = 0; q2 < 32; ++q2) {
			const double a = -2.0 * M_PI * (double)(j * q2) / 512.0;
			tw.push_back (make_float2 ((float)std::cos (a), (float)std::sin (a)));
		}
	}
	return tw;
}

} // namespace prk
