#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > $O/s12_tests.log 2>&1; tail -4 $O/s12_tests.log
timeout 300 python scripts/ntt_probe2.py 27 3 > $O/s12_ntt.log 2>&1; tail -2 $O/s12_ntt.log
