#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > $O/s20_tests.log 2>&1; tail -3 $O/s20_tests.log
export PROBE_CHECK=1 BZ_MSM_PRECOMP=2
timeout 300 python scripts/perf_probe.py 23,26 0 > $O/s20_probe.log 2>&1; grep logn $O/s20_probe.log | cut -c1-330
