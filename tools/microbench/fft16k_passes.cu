// Microbenchmark: cycles per pass of the fft16k.cuh segment transform, each pass
// looped in isolation on one CTA per SM (148 CTAs x 512 threads, 133 KB smem),
// to separate FMA-pipe time, shared-memory pipe time and their overlap.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o fft16k_passes fft16k_passes.cu
#include <cstdio>
#include <vector>

#include "../../phaserotate/lv2_b200/csrc/fft16k_tables.h"
#include "../../phaserotate/lv2_b200/csrc/kernels.cuh"

using namespace prk;

#ifndef STAGGER
#define STAGGER 0
#endif
#ifndef DELAY_NS
#define DELAY_NS 250
#endif

template <int WHICH>
__global__ void __launch_bounds__ (512, 1) pass_kernel (const float4* src4, const float2* src2, const float2* tw, const float4* G, float2* out, int iters, long long* cyc)
{
	extern __shared__ __align__ (16) float2 sm[];
	const int tid = threadIdx.x;
	for (int i = tid; i < kM; i += 512) sm[i] = make_float2 (1e-3f * (i & 255), -1e-3f * (i & 127));
	__syncthreads ();
	const long long t0 = clock64 ();
	float2          acc = make_float2 (0.f, 0.f);
	for (int it = 0; it < iters; ++it) {
		if (WHICH == 0) p1_forward (sm, tw, tid, Inter1Loader { src2 + (size_t)(blockIdx.x * 7 + it) % 64 * kM });
		if (WHICH == 1) p1_forward (sm, tw, tid, Inter2Loader { src4 + (size_t)(blockIdx.x * 7 + it) % 64 * kM, it & 1 });
		if (WHICH == 2) p2_pass<-1> (sm, tid);
		if (WHICH == 3) mid_pass (sm, GTable { G, tid, 0 }, tw + kTwP1Rows * 512, tid);
		if (WHICH == 8) { // MID, upper half of the warps delayed
			if (tid >= 256) __nanosleep (DELAY_NS);
			mid_pass (sm, GTable { G, tid, 0 }, tw + kTwP1Rows * 512, tid);
		}
		if (WHICH == 9) { // P2, upper half of the warps delayed
			if (tid >= 256) __nanosleep (DELAY_NS);
			p2_pass<-1> (sm, tid);
		}
		if (WHICH == 10) { // MID, quarter offsets
			if (tid >= 128) __nanosleep (DELAY_NS / 2 * (tid >> 7));
			mid_pass (sm, GTable { G, tid, 0 }, tw + kTwP1Rows * 512, tid);
		}
		if (WHICH == 4) {
			float2 w[32];
			p1_inverse (sm, tw, tid, w);
#pragma unroll
			for (int k = 0; k < 32; ++k) acc = cadd (acc, w[k]);
		}
		if (WHICH == 5) { // butterfly only, operands stay in registers
			float2 u[32];
#pragma unroll
			for (int k = 0; k < 32; ++k) u[k] = make_float2 (acc.x + k, acc.y - k);
			dft32<-1> (u);
#pragma unroll
			for (int k = 0; k < 32; ++k) acc = cadd (acc, u[k]);
		}
		if (WHICH == 6) { // shared-memory traffic of P2 only
			const int q1 = tid >> 4, j = tid & 15;
			float2    u[32];
#pragma unroll
			for (int k = 0; k < 32; ++k) u[k] = sm[swz (q1 * 32 + k, j)];
#pragma unroll
			for (int q = 0; q < 32; ++q) sm[swz (q1 * 32 + q, j)] = u[(q + 1) & 31];
		}
		if (WHICH == 7) { // 16-point butterflies + twiddles of MID only
			float2 u[16];
#pragma unroll
			for (int k = 0; k < 16; ++k) u[k] = make_float2 (acc.x + k, acc.y - k);
#pragma unroll
			for (int r = 0; r < 2; ++r) {
#pragma unroll
				for (int j = 1; j < 16; ++j) u[j] = cmul (u[j], make_float2 (acc.y, acc.x));
				dft16<-1> (u);
#pragma unroll
				for (int j = 0; j < 16; ++j) u[j] = cmul (u[j], make_float2 (acc.x, acc.y));
				dft16<+1> (u);
#pragma unroll
				for (int j = 1; j < 16; ++j) u[j] = cmulc (u[j], make_float2 (acc.y, acc.x));
			}
#pragma unroll
			for (int k = 0; k < 16; ++k) acc = cadd (acc, u[k]);
		}
		__syncthreads ();
	}
	const long long t1 = clock64 ();
	if (tid == 0) cyc[blockIdx.x] = t1 - t0;
	out[blockIdx.x * 512 + tid] = cadd (acc, sm[tid]);
}

template <int WHICH>
void run (const char* name, const float4* src4, const float2* tw, const float4* G, float2* out, long long* cyc)
{
	const int smem = kSmemBytes, iters = 64;
	cudaFuncSetAttribute (pass_kernel<WHICH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	pass_kernel<WHICH><<<148, 512, smem>>> (src4, (const float2*)src4, tw, G, out, 4, cyc);
	pass_kernel<WHICH><<<148, 512, smem>>> (src4, (const float2*)src4, tw, G, out, iters, cyc);
	cudaError_t e = cudaDeviceSynchronize ();
	long long   h[148];
	cudaMemcpy (h, cyc, sizeof (h), cudaMemcpyDeviceToHost);
	double avg = 0;
	for (int i = 0; i < 148; ++i) avg += h[i];
	printf ("%-28s %8.0f cycles/pass  (%s)\n", name, avg / 148 / iters, cudaGetErrorString (e));
}

int main ()
{
	float4* src4;
	cudaMalloc (&src4, sizeof (float4) * kM * 65);
	cudaMemset (src4, 0, sizeof (float4) * kM * 65);
	std::vector<float>  g (4096, 0.01f);
	std::vector<float2> G  = make_filter_spectrum (g.data (), 4096);
	std::vector<float2> tw = make_twiddles ();
	float2 *dG, *dtw, *out;
	long long* cyc;
	cudaMalloc (&dG, sizeof (float2) * G.size ());
	cudaMalloc (&dtw, sizeof (float2) * tw.size ());
	cudaMalloc (&out, sizeof (float2) * 148 * 512);
	cudaMalloc (&cyc, sizeof (long long) * 148);
	cudaMemcpy (dG, G.data (), sizeof (float2) * G.size (), cudaMemcpyHostToDevice);
	cudaMemcpy (dtw, tw.data (), sizeof (float2) * tw.size (), cudaMemcpyHostToDevice);
	printf ("STAGGER=%d\n", STAGGER);
	run<0> ("P1 forward (mono, L2 hits)", src4, dtw, (const float4*)dG, out, cyc);
	run<1> ("P1 forward (stereo, L2 hits)", src4, dtw, (const float4*)dG, out, cyc);
	run<2> ("P2", src4, dtw, (const float4*)dG, out, cyc);
	run<3> ("MID", src4, dtw, (const float4*)dG, out, cyc);
	run<4> ("P1 inverse (to registers)", src4, dtw, (const float4*)dG, out, cyc);
	run<5> ("dft32 only (+64 packed adds)", src4, dtw, (const float4*)dG, out, cyc);
	run<6> ("P2 smem traffic only", src4, dtw, (const float4*)dG, out, cyc);
	run<7> ("MID arithmetic only (2 rows)", src4, dtw, (const float4*)dG, out, cyc);
	printf ("DELAY_NS=%d\n", DELAY_NS);
	run<8> ("MID, half the warps delayed", src4, dtw, (const float4*)dG, out, cyc);
	run<9> ("P2, half the warps delayed", src4, dtw, (const float4*)dG, out, cyc);
	run<10> ("MID, quarter offsets", src4, dtw, (const float4*)dG, out, cyc);
	return 0;
}
