#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > $O/s6_tests.log 2>&1; tail -5 $O/s6_tests.log
(time timeout 900 python bench.py) > $O/s6_bench.log 2>&1; tail -2 $O/s6_bench.log | cut -c1-1200
export PROBE_CHECK=0 BZ_MSM_PRECOMP=2
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/s6_launches_merged_2p26.csv python scripts/perf_probe.py 26 0 > $O/s6_probe26.log 2>&1
tail -1 $O/s6_probe26.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_accumulate -c 1 -f -o $O/s6_acc_merged_2p22 python scripts/perf_probe.py 22 0 > $O/s6_ncu_acc.log 2>&1
tail -2 $O/s6_ncu_acc.log
ls -la $O/*.ncu-rep
