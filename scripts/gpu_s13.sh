#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
export PROBE_CHECK=1 BZ_MSM_PRECOMP=2
timeout 300 python scripts/perf_probe.py 24 0 > $O/s13_probe_a.log 2>&1; tail -1 $O/s13_probe_a.log | cut -c1-420
BLAZE_B200_LIB=$PWD/blaze_b200/libblaze_b200_kara.so timeout 300 python scripts/perf_probe.py 24 0 > $O/s13_probe_kara.log 2>&1; tail -1 $O/s13_probe_kara.log | cut -c1-420
