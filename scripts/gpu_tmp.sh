#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > $O/s24_tests.log 2>&1; tail -3 $O/s24_tests.log
(time timeout 900 python bench.py) > $O/s24_bench.log 2>&1; tail -2 $O/s24_bench.log | cut -c1-200
