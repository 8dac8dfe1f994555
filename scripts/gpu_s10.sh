#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
nvidia-smi -L
(time timeout 600 python -m pytest tests/test_ntt_gpu.py -m gpu -x -q) > $O/s10_tests.log 2>&1; tail -4 $O/s10_tests.log
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3) > $O/s10_bench2.log 2>&1; tail -3 $O/s10_bench2.log | cut -c1-2500
