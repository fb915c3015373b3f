// Microbenchmark: issue rate of packed fp32 (FFMA2/FADD2/FMUL2, sm_100) against
// scalar FFMA/FADD, alone and mixed with ALU-pipe and LDS work.  Prints cycles per
// loop iteration per warp scheduler so the rates can be read against the SASS body
// (cuobjdump -sass).  Used to decide how the FFT passes of fftconv_kernel are written.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define NCH 8
constexpr int kIters = 4096;

template <int MODE>
__global__ void __launch_bounds__ (512) bench (float2* out, float2 a, float2 b, int iters, long long* cyc)
{
	__shared__ float2 smem[2048];
	float2 x[NCH];
	int    ii[NCH];
#pragma unroll
	for (int i = 0; i < NCH; ++i) {
		x[i]  = make_float2 (threadIdx.x * 0.001f + i, threadIdx.x * 0.002f - i);
		ii[i] = threadIdx.x * 7 + i;
	}
	for (int i = threadIdx.x; i < 2048; i += blockDim.x) smem[i] = make_float2 (i, -i);
	__syncthreads ();
	const float2* sp = smem + (threadIdx.x & 1023);
	const long long t0 = clock64 ();
#pragma unroll 1
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int i = 0; i < NCH; ++i) {
			if (MODE == 0) { // scalar FFMA x2 (one complex)
				x[i].x = fmaf (x[i].x, a.x, b.x);
				x[i].y = fmaf (x[i].y, a.y, b.y);
			} else if (MODE == 1) { // FFMA2
				x[i] = __ffma2_rn (x[i], a, b);
			} else if (MODE == 2) { // scalar FADD x2
				x[i].x = x[i].x + a.x;
				x[i].y = x[i].y + a.y;
			} else if (MODE == 3) { // FADD2
				x[i] = __fadd2_rn (x[i], a);
			} else if (MODE == 4) { // FMUL2
				x[i] = __fmul2_rn (x[i], a);
			} else if (MODE == 5) { // FFMA2 with broadcast operand (register)
				x[i] = __ffma2_rn (make_float2 (x[(i + 1) % NCH].x, x[(i + 1) % NCH].x), a, x[i]);
			} else if (MODE == 6) { // FFMA2 + 1 ALU op per FFMA2
				x[i]  = __ffma2_rn (x[i], a, b);
				ii[i] = (ii[i] ^ it) + 0x1234567;
			} else if (MODE == 7) { // scalar FFMA pair + 1 ALU op
				x[i].x = fmaf (x[i].x, a.x, b.x);
				x[i].y = fmaf (x[i].y, a.y, b.y);
				ii[i]  = (ii[i] ^ it) + 0x1234567;
			} else if (MODE == 8) { // FFMA2 + 2 ALU ops
				x[i]  = __ffma2_rn (x[i], a, b);
				ii[i] = ((ii[i] ^ it) + 0x1234567) ^ (ii[i] >> 3);
			} else if (MODE == 9) { // FADD2 + LDS.64 per 2 FADD2
				x[i] = __fadd2_rn (x[i], a);
				if (i & 1) x[i] = __fadd2_rn (x[i], sp[(i * 64 + it) & 1023]);
			} else if (MODE == 10) { // complex multiply, packed: 2 instr
				const float2 t = __fmul2_rn (make_float2 (x[i].y, x[i].y), b);
				x[i]           = __ffma2_rn (make_float2 (x[i].x, x[i].x), a, t);
			} else if (MODE == 11) { // complex multiply, scalar: 4 instr
				const float re = fmaf (x[i].x, a.x, -x[i].y * a.y);
				const float im = fmaf (x[i].x, a.y, x[i].y * a.x);
				x[i]           = make_float2 (re, im);
			} else if (MODE == 12) { // only ALU: 2 ops
				ii[i] = ((ii[i] ^ it) + 0x1234567) ^ (ii[i] >> 3);
			}
		}
	}
	const long long t1 = clock64 ();
	float2 s = make_float2 (0.f, 0.f);
	int    si = 0;
#pragma unroll
	for (int i = 0; i < NCH; ++i) {
		s.x += x[i].x;
		s.y += x[i].y;
		si += ii[i];
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = make_float2 (s.x + si, s.y);
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run (const char* name, int threads)
{
	const int grid = 148;
	float2*   out;
	long long* cyc;
	cudaMalloc (&out, sizeof (float2) * grid * 1024);
	cudaMalloc (&cyc, sizeof (long long) * grid);
	const float2 a = make_float2 (1.0000001f, 0.9999999f), b = make_float2 (1e-9f, -1e-9f);
	bench<MODE><<<grid, threads>>> (out, a, b, 16, cyc);
	cudaEvent_t e0, e1;
	cudaEventCreate (&e0);
	cudaEventCreate (&e1);
	cudaEventRecord (e0);
	bench<MODE><<<grid, threads>>> (out, a, b, kIters, cyc);
	cudaEventRecord (e1);
	cudaDeviceSynchronize ();
	float ms;
	cudaEventElapsedTime (&ms, e0, e1);
	long long h[148];
	cudaMemcpy (h, cyc, sizeof (h), cudaMemcpyDeviceToHost);
	double avg = 0;
	for (int i = 0; i < grid; ++i) avg += h[i];
	avg /= grid;
	const int warps_per_smsp = threads / 32 / 4;
	// cycles per (iteration of one warp) per scheduler = cycles / (iters * warps on that scheduler)
	printf ("%-44s threads %4d  %8.3f ms  cyc/iter/warp-on-smsp %7.3f  (per chain-step %6.3f)\n", name, threads, ms,
	        avg / kIters / warps_per_smsp, avg / kIters / warps_per_smsp / NCH);
	cudaFree (out);
	cudaFree (cyc);
}

int main ()
{
	for (int threads : { 256, 512, 1024 }) {
		run<0> ("0 scalar FFMA x2 per step", threads);
		run<1> ("1 FFMA2 per step", threads);
		run<2> ("2 scalar FADD x2 per step", threads);
		run<3> ("3 FADD2 per step", threads);
		run<4> ("4 FMUL2 per step", threads);
		run<5> ("5 FFMA2 bcast operand per step", threads);
		run<6> ("6 FFMA2 + 1 ALU per step", threads);
		run<7> ("7 scalar FFMA x2 + 1 ALU per step", threads);
		run<8> ("8 FFMA2 + ~3 ALU per step", threads);
		run<9> ("9 FADD2 (+ LDS.64 & FADD2 every 2nd)", threads);
		run<10> ("10 packed cmul (FMUL2+FFMA2)", threads);
		run<11> ("11 scalar cmul (2 FMUL + 2 FFMA)", threads);
		run<12> ("12 ALU only (~3 ops)", threads);
	}
	return 0;
}
