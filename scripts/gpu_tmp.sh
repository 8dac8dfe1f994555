#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > $O/s22_tests.log 2>&1; tail -3 $O/s22_tests.log
(time timeout 900 python bench.py) > $O/s22_bench.log 2>&1; tail -2 $O/s22_bench.log | cut -c1-200
export PROBE_CHECK=0 BZ_MSM_PRECOMP=2
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/s22_launches_merged_2p26.csv python scripts/perf_probe.py 26 0 > $O/s22_probe26.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/s22_launches_ntt_2p27.csv python scripts/ntt_probe2.py 27 2 > $O/s22_ntt.log 2>&1
python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-200
