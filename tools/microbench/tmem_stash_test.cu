// Checks the thread-private TMEM stash (tmem_stash.cuh): every thread of a
// 512-thread CTA stores 32 complex values, all CTAs sync, and reads them back
// in shifted groups of four like the fftconv epilogue does.
#include <cstdio>
#include "../../phaserotate/lv2_b200/csrc/tmem_stash.cuh"
using namespace prk;

__global__ void __launch_bounds__ (512, 1) k (int* bad, long long* cyc, int iters)
{
	__shared__ uint32_t slot;
	const int      tid  = threadIdx.x;
	const uint32_t base = tmem_alloc_all (&slot, tid);
	const uint32_t tb   = tmem_thread_base (base, tid);
	int            nbad = 0;
	long long      t0   = clock64 ();
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int g = 0; g < 8; ++g) {
			float2 v[4];
#pragma unroll
			for (int j = 0; j < 4; ++j) v[j] = make_float2 ((float)(tid * 64 + (4 * g + j) * 2 + it), (float)(tid * 64 + (4 * g + j) * 2 + 1 + it + blockIdx.x));
			tmem_st4 (tb + 8 * g, v[0], v[1], v[2], v[3]);
		}
		tmem_wait_st ();
		__syncthreads ();
		const int dk = 4;
#pragma unroll
		for (int kb = 8; kb < 32; kb += 4) {
			float2 v[4];
			tmem_ld4 (tb + 2 * (kb - dk), v);
#pragma unroll
			for (int j = 0; j < 4; ++j) {
				const int k = kb - dk + j;
				if (v[j].x != (float)(tid * 64 + k * 2 + it) || v[j].y != (float)(tid * 64 + k * 2 + 1 + it + blockIdx.x)) ++nbad;
			}
		}
		__syncthreads ();
	}
	long long t1 = clock64 ();
	if (nbad) atomicAdd (bad, nbad);
	if (tid == 0) cyc[blockIdx.x] = t1 - t0;
	tmem_free_all (base, tid);
}

int main ()
{
	int* bad; long long* cyc;
	cudaMalloc (&bad, 4); cudaMemset (bad, 0, 4);
	cudaMalloc (&cyc, 8 * 148);
	k<<<148, 512>>> (bad, cyc, 4);
	k<<<148, 512>>> (bad, cyc, 64);
	cudaError_t e = cudaDeviceSynchronize ();
	int h; long long hc[148];
	cudaMemcpy (&h, bad, 4, cudaMemcpyDeviceToHost);
	cudaMemcpy (hc, cyc, sizeof (hc), cudaMemcpyDeviceToHost);
	printf ("tmem stash: %s, mismatches %d, %.0f cycles per (store 32 + load 24) round\n", cudaGetErrorString (e), h, (double)hc[0] / 64);
	return h != 0 || e != cudaSuccess;
}
