#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
(time timeout 900 python -m pytest tests/test_ntt_gpu.py -m gpu -x -q) > $O/s21_tests.log 2>&1; tail -3 $O/s21_tests.log
free -g | head -2
(time timeout 900 python bench.py --no-dma --no-cpu-baseline --steps 3) > $O/s21_bench.log 2>&1; tail -2 $O/s21_bench.log | cut -c1-200
