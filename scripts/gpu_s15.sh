#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > $O/s15_tests.log 2>&1; tail -4 $O/s15_tests.log
(time timeout 900 python bench.py --no-ntt --no-cpu-baseline --steps 3) > $O/s15_bench.log 2>&1; tail -2 $O/s15_bench.log | cut -c1-300
