#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
(time timeout 900 python -m pytest tests/test_msm_gpu.py -m gpu -x -q) > $O/s9_tests.log 2>&1; tail -4 $O/s9_tests.log
export PROBE_CHECK=1 BZ_MSM_PRECOMP=2
timeout 300 python scripts/perf_probe.py 24 0 > $O/s9_probe_a.log 2>&1; tail -1 $O/s9_probe_a.log
BZ_MSM_TMA=1 timeout 300 python scripts/perf_probe.py 24 0 > $O/s9_probe_tma.log 2>&1; tail -1 $O/s9_probe_tma.log
export PROBE_CHECK=0
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/s9_launches_merged_2p23.csv python scripts/perf_probe.py 23 0 > $O/s9_probe23.log 2>&1
tail -1 $O/s9_probe23.log
