#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
export PROBE_CHECK=0 BZ_MSM_PRECOMP=2
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/s4_launches_merged_2p26.csv python scripts/perf_probe.py 26 0 > $O/s4_probe.log 2>&1
tail -2 $O/s4_probe.log
BZ_MSM_PRECOMP=2 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/s4_launches_merged_2p24_c23.csv python scripts/perf_probe.py 24 23 > $O/s4_probe23.log 2>&1
