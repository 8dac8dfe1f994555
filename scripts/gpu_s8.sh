#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
export PROBE_CHECK=1 BZ_MSM_PRECOMP=2
BLAZE_B200_LIB=$PWD/blaze_b200/libblaze_b200_mb4.so timeout 300 python scripts/perf_probe.py 24 0 > $O/s8_probe_mb4.log 2>&1; tail -1 $O/s8_probe_mb4.log
unset BZ_MSM_PRECOMP
(time timeout 900 python -m pytest tests -m gpu -x -q) > $O/s8_tests.log 2>&1; tail -4 $O/s8_tests.log
(time timeout 900 python bench.py) > $O/s8_bench.log 2>&1; tail -2 $O/s8_bench.log | cut -c1-600
timeout 600 python bench.py --curve BN254 --log-n 24 --no-ntt --no-cpu-baseline > $O/s8_bench_bn254.log 2>&1; tail -1 $O/s8_bench_bn254.log | cut -c1-900
timeout 600 python bench.py --curve BLS377 --no-ntt --no-cpu-baseline > $O/s8_bench_377.log 2>&1; tail -1 $O/s8_bench_377.log | cut -c1-900
