#!/bin/bash
# ncu passes of one round (run on the GPU box through gpurun): launch list of our kernels
# for the bench command, DRAM traffic of every FFT-convolution launch, one full capture.
# Usage: tools/ncu_round.sh <tag>      -> gpurun_out/ncu_<tag>_*.csv / .ncu-rep
set -u
TAG=${1:-vX}
mkdir -p gpurun_out
K='regex:fftconv_kernel|sweep_kernel|sweep_window_kernel|sector_thr_kernel|threshold_kernel|truepeak_kernel|tp_carry|interleave_kernel|fir_stream|peer_max|pcm'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv \
    --log-file gpurun_out/ncu_${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extra --no-strong --no-config5 > gpurun_out/ncu_${TAG}_bench_under_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k 'regex:fftconv_kernel' -s 6 -c 6 --csv \
    --log-file gpurun_out/ncu_${TAG}_traffic.csv python tools/prof_sweep.py --seconds 3600 --steps 4 > gpurun_out/ncu_${TAG}_traffic.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:fftconv_kernel' -s 5 -c 1 -o gpurun_out/ncu_${TAG}_conv -f \
    python tools/prof_sweep.py --seconds 3600 --steps 2 > gpurun_out/ncu_${TAG}_full.log 2>&1
ls -la gpurun_out | tail -8
