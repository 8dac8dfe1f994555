#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
export PROBE_CHECK=1 BZ_MSM_PRECOMP=2 PROBE_CURVE=BN254
timeout 300 python scripts/perf_probe.py 24 0 2>&1 | grep logn | cut -c1-330
BLAZE_B200_LIB=$PWD/blaze_b200/libblaze_b200_n8.so timeout 300 python scripts/perf_probe.py 24 0 2>&1 | grep logn | cut -c1-330
