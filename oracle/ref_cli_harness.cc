/*
 * ORACLE / TEST INFRASTRUCTURE — not part of the product.
 *
 * Wraps the UNMODIFIED reference CLI translation unit
 * (/root/reference/cli/phase-rotate.cc, found through -I; never copied into
 * this repository) so that tests and the CPU-baseline leg of bench.py can call
 * its classes directly and read `_peak[c][a]` at full float precision:
 *
 *   ref_cli_analyze  -> reference analyze_file()          (cli/phase-rotate.cc:565-587)
 *   ref_cli_apply    -> reference PhaseRotate::apply()    (cli/phase-rotate.cc:467-485)
 *   ref_cli_lut      -> reference SinCosLut               (cli/phase-rotate.cc:41-74)
 *
 * Built by oracle/Makefile into oracle/_ref/libref_cli.so against the stand-in
 * fftw3/sndfile headers in standin/ (the real libraries are absent here).
 */
#define main phase_rotate_reference_main
#include "phase-rotate.cc"
#undef main

#include <chrono>

extern "C" {

/* Runs the reference's whole-file analysis pass on in-memory audio.
 * peaks_out: [n_chn][MAXSAMPLE] as left in PhaseRotate::_peak after the pass.
 * Returns the wall time of analyze_file() in seconds, or < 0 on error. */
double
ref_cli_analyze (const float* interleaved, int64_t n_frames, int n_chn, int blksiz,
                 int ang_start, int ang_end, int ang_stride, int only_chn, float* peaks_out)
{
	SF_INFO  nfo;
	SNDFILE* sf = standin_sf_open_memory (interleaved, n_frames, n_chn, 48000, &nfo);
	if (!sf) {
		return -1;
	}
	float* buf = (float*)malloc ((size_t)blksiz * n_chn * sizeof (float));
	double dt  = -1;
	{
		PRPVec prp;
		for (int i = 0; i < n_chn; ++i) {
			prp.push_back (std::unique_ptr<PhaseRotateProc> (new PhaseRotateProc (blksiz)));
		}
		PhaseRotate pr (prp, n_chn);
		auto        t0 = std::chrono::steady_clock::now ();
		analyze_file (pr, sf, buf, ang_start, ang_end, ang_stride, only_chn);
		auto t1 = std::chrono::steady_clock::now ();
		dt      = std::chrono::duration<double> (t1 - t0).count ();
		if (peaks_out) {
			for (int c = 0; c < n_chn; ++c) {
				for (int a = 0; a < MAXSAMPLE; ++a) {
					peaks_out[c * MAXSAMPLE + a] = pr.peak (c, a);
				}
			}
		}
	}
	free (buf);
	sf_close (sf);
	return dt;
}

/* Streams ceil(n_frames / blksiz) zero-padded blocks plus `flush_blocks`
 * all-zero blocks through PhaseRotate::apply and returns every processed block
 * back to back in out[(nblk + flush_blocks) * blksiz * n_chn] (interleaved, no
 * latency trim — the trim lives in main()'s write loop and is exercised
 * through the oracle/_ref/phase-rotate binary instead). */
double
ref_cli_apply (const float* interleaved, int64_t n_frames, int n_chn, int blksiz,
               const int* angles, int flush_blocks, float* out)
{
	PRPVec prp;
	for (int i = 0; i < n_chn; ++i) {
		prp.push_back (std::unique_ptr<PhaseRotateProc> (new PhaseRotateProc (blksiz)));
	}
	PhaseRotate      pr (prp, n_chn);
	std::vector<int> ang (angles, angles + n_chn);
	const size_t     bs   = (size_t)blksiz * n_chn;
	const int64_t    nblk = (n_frames + blksiz - 1) / blksiz;
	auto             t0   = std::chrono::steady_clock::now ();
	for (int64_t b = 0; b < nblk + flush_blocks; ++b) {
		float* dst = out + (size_t)b * bs;
		memset (dst, 0, bs * sizeof (float));
		if (b < nblk) {
			const int64_t f0 = b * blksiz;
			const int64_t n  = std::min<int64_t> (blksiz, n_frames - f0);
			memcpy (dst, interleaved + (size_t)f0 * n_chn, (size_t)n * n_chn * sizeof (float));
		}
		pr.apply (dst, ang);
	}
	auto t1 = std::chrono::steady_clock::now ();
	return std::chrono::duration<double> (t1 - t0).count ();
}

void
ref_cli_lut (float* s_out, float* c_out)
{
	for (int a = 0; a < MAXSAMPLE; ++a) {
		scl.sincos (a, &s_out[a], &c_out[a]);
	}
}

int
ref_cli_maxsample (void)
{
	return MAXSAMPLE;
}

/* Hilbert FIR taps as the reference designs them (constructor,
 * cli/phase-rotate.cc:144-161), recovered by pushing a unit impulse through
 * PhaseRotateProc::hilbert: taps[0..blksiz). */
void
ref_cli_taps (int blksiz, float* taps)
{
	PhaseRotateProc    p (blksiz);
	std::vector<float> tdc (2 * blksiz, 0.f), olp (blksiz, 0.f);
	tdc[blksiz] = 1.f;
	p.hilbert (tdc.data (), taps, olp.data ());
}

} /* extern "C" */
