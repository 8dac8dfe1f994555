#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
(time timeout 900 python -m pytest tests/test_msm_gpu.py -m gpu -x -q) > $O/s5_tests.log 2>&1; tail -15 $O/s5_tests.log
export PROBE_CHECK=1
BZ_MSM_PRECOMP=0 timeout 300 python scripts/perf_probe.py 24 0 > $O/s5_probe_plain.log 2>&1; tail -1 $O/s5_probe_plain.log
BZ_MSM_PRECOMP=2 timeout 300 python scripts/perf_probe.py 24 0 > $O/s5_probe_merged.log 2>&1; tail -1 $O/s5_probe_merged.log
export PROBE_CHECK=0 BZ_MSM_PRECOMP=2
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/s5_launches_merged_2p26.csv python scripts/perf_probe.py 26 0 > $O/s5_probe26.log 2>&1
tail -1 $O/s5_probe26.log
