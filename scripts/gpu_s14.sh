#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
(time timeout 900 python -m pytest tests/test_msm_gpu.py -m gpu -x -q) > $O/s14_tests.log 2>&1; tail -4 $O/s14_tests.log
(time timeout 900 python bench.py) > $O/s14_bench.log 2>&1; tail -2 $O/s14_bench.log | cut -c1-300
export PROBE_CHECK=0 BZ_MSM_PRECOMP=2
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_accumulate -c 1 -f -o $O/s14_acc_3cta_2p22 python scripts/perf_probe.py 22 0 > $O/s14_ncu_acc.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_part_scatter -c 2 -f -o $O/s14_scatter_2p24 python scripts/perf_probe.py 24 0 > $O/s14_ncu_scatter.log 2>&1
ls -la $O/*.ncu-rep
