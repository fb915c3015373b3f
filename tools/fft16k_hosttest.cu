// Host emulation of the fft16k.cuh passes (scalar arithmetic, thread loop in
// place of the CTA): checks the index algebra, swizzle, twiddle and filter
// tables of the GPU FFT convolution against a direct convolution in double.
// No GPU needed.  Build: nvcc -O2 -o fft16k_hosttest tools/fft16k_hosttest.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../phaserotate/lv2_b200/csrc/fft16k_tables.h"

using namespace prk;

struct VecLoader {
	const float2* z;
	struct T {
		const float2* p;
		float2 operator() (int off) const { return p[off]; }
	};
	T thread (int e) const { return T { z + e }; }
};

int
main (int argc, char** argv)
{
	const int Lh = argc > 1 ? atoi (argv[1]) : 4096;
	srand (12345);
	std::vector<float>  g ((size_t)Lh);
	for (auto& v : g) v = (float)rand () / RAND_MAX - 0.5f;
	std::vector<float2> z ((size_t)kM);
	for (auto& v : z) v = make_float2 ((float)rand () / RAND_MAX - 0.5f, (float)rand () / RAND_MAX - 0.5f);

	const std::vector<float2> G  = make_filter_spectrum (g.data (), Lh);
	const std::vector<float2> tw = make_twiddles ();
	const float2* twp1 = tw.data ();
	const float2* twm  = tw.data () + kTwP1Rows * 512;

	// bank-conflict check of the swizzle: every half-warp of P1/P2 accesses and
	// every quarter-warp of MID accesses must hit distinct 16-byte... (8-byte for P1/P2) slots mod 128 bytes
	int conflicts = 0;
	for (int q = 0; q < 32; ++q) {
		for (int hw = 0; hw < 32; ++hw) { // half-warps of P1: e = 16 hw .. 16 hw + 15
			unsigned seen = 0;
			for (int l = 0; l < 16; ++l) {
				const int e = 16 * hw + l;
				const int p = swz (q * 32 + (e >> 4), e & 15) & 15;
				if (seen & (1u << p)) ++conflicts;
				seen |= 1u << p;
			}
		}
	}
	for (int c = 0; c < 8; ++c) {
		for (int qw = 0; qw < 128; ++qw) { // quarter-warps of MID: rows 8 qw .. 8 qw + 7
			unsigned seen = 0;
			for (int l = 0; l < 8; ++l) {
				const int row = 8 * qw + l;
				const int s   = ((row >> 5) ^ row) & 7;
				const int p   = c ^ s;
				if (seen & (1u << p)) ++conflicts;
				seen |= 1u << p;
			}
		}
	}
	// swz must be a permutation
	{
		std::vector<char> hit ((size_t)kM, 0);
		for (int i = 0; i < kM; ++i) {
			const int p = swz (i >> 4, i & 15);
			if (p < 0 || p >= kM || hit[(size_t)p]) ++conflicts;
			else hit[(size_t)p] = 1;
		}
	}

	std::vector<float2> sm ((size_t)kM);
	if (argc > 2) {
		// two tap partitions (FIR length 32768): Lh = 2 * Lp taps, Lp = kM / 2, stream of kM + Lp points;
		// segment A = s[0, kM) leaves its spectrum in the scratch, segment B = s[Lp, Lp + kM) is convolved
		const int Lp = kM / 2, L2 = 2 * Lp;
		std::vector<float> g2 ((size_t)L2);
		for (auto& v : g2) v = (float)rand () / RAND_MAX - 0.5f;
		std::vector<float2> s ((size_t)(kM + Lp));
		for (auto& v : s) v = make_float2 ((float)rand () / RAND_MAX - 0.5f, (float)rand () / RAND_MAX - 0.5f);
		const std::vector<float2> G0 = make_filter_spectrum (g2.data (), Lp), G1 = make_filter_spectrum (g2.data () + Lp, Lp);
		std::vector<float4> scr ((size_t)kM / 2);
		for (int e = 0; e < kConvThreads; ++e) p1_forward (sm.data (), twp1, e, VecLoader { s.data () });
		for (int t = 0; t < kConvThreads; ++t) p2_pass<-1> (sm.data (), t);
		for (int t = 0; t < kConvThreads; ++t) mid_pass<MID_SPECTRUM> (sm.data (), GTable { nullptr, 0, 0 }, twm, t, scr.data () + t);
		for (int e = 0; e < kConvThreads; ++e) p1_forward (sm.data (), twp1, e, VecLoader { s.data () + Lp });
		for (int t = 0; t < kConvThreads; ++t) p2_pass<-1> (sm.data (), t);
		for (int t = 0; t < kConvThreads; ++t) {
			mid_pass<MID_CONV2> (sm.data (), GTable { nullptr, 0, 0 }, twm, t, scr.data () + t, reinterpret_cast<const float4*> (G0.data ()),
			                     reinterpret_cast<const float4*> (G1.data ()));
		}
		for (int t = 0; t < kConvThreads; ++t) p2_pass<+1> (sm.data (), t);
		double maxerr = 0.0, maxref = 0.0;
		for (int e = 0; e < kConvThreads; ++e) {
			float2 w[32];
			p1_inverse (sm.data (), twp1, e, w);
			for (int k = 16; k < 32; ++k) {
				const int i = e + 512 * k; // valid outputs of B: i >= Lp, stream index Lp + i
				if ((i % 41) != 0) continue;
				double re = 0.0, im = 0.0;
				for (int j = 0; j < L2; ++j) {
					re += (double)g2[(size_t)j] * s[(size_t)(Lp + i - j)].x;
					im += (double)g2[(size_t)j] * s[(size_t)(Lp + i - j)].y;
				}
				maxerr = std::max (maxerr, std::max (std::fabs (re - w[k].x), std::fabs (im - w[k].y)));
				maxref = std::max (maxref, std::max (std::fabs (re), std::fabs (im)));
			}
		}
		printf ("two partitions  max abs err %.3e  max |ref| %.3e  rel %.3e\n", maxerr, maxref, maxerr / maxref);
		return (maxerr / maxref < 2e-6 && conflicts == 0) ? 0 : 1;
	}
	for (int e = 0; e < kConvThreads; ++e) p1_forward (sm.data (), twp1, e, VecLoader { z.data () });
	for (int t = 0; t < kConvThreads; ++t) p2_pass<-1> (sm.data (), t);
	for (int t = 0; t < kConvThreads; ++t) mid_pass (sm.data (), GTable { reinterpret_cast<const float4*> (G.data ()), t, 0 }, twm, t);
	for (int t = 0; t < kConvThreads; ++t) p2_pass<+1> (sm.data (), t);
	std::vector<float2> out ((size_t)kM);
	for (int e = 0; e < kConvThreads; ++e) {
		float2 w[32];
		p1_inverse (sm.data (), twp1, e, w);
		for (int k = 0; k < 32; ++k) out[(size_t)(e + 512 * k)] = w[k];
	}

	// direct linear convolution for the valid outputs i >= Lh - 1 (circular == linear there)
	double maxerr = 0.0, maxref = 0.0;
	for (int i = Lh - 1; i < kM; i += 37) {
		double re = 0.0, im = 0.0;
		for (int j = 0; j < Lh; ++j) {
			re += (double)g[(size_t)j] * z[(size_t)(i - j)].x;
			im += (double)g[(size_t)j] * z[(size_t)(i - j)].y;
		}
		maxerr = std::max (maxerr, std::max (std::fabs (re - out[(size_t)i].x), std::fabs (im - out[(size_t)i].y)));
		maxref = std::max (maxref, std::max (std::fabs (re), std::fabs (im)));
	}
	printf ("Lh %d  max abs err %.3e  max |ref| %.3e  rel %.3e  swizzle conflicts %d\n", Lh, maxerr, maxref, maxerr / maxref, conflicts);
	return (maxerr / maxref < 2e-6 && conflicts == 0) ? 0 : 1;
}
