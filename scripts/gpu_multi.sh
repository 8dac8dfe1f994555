#!/bin/bash
# multi-GPU validation: bash scripts/gpu_multi.sh N  (bench at N ranks + distributed NTT parity at 2^21 and 2^24)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${1:-2}
O=gpurun_out
nvidia-smi -L | wc -l
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 8 --warmup 3) > $O/multi${N}_bench.log 2>&1; tail -2 $O/multi${N}_bench.log | cut -c1-400
for L in 21 24; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/dist_ntt_check.py $L > $O/multi${N}_ntt$L.log 2>&1; grep DIST_NTT $O/multi${N}_ntt$L.log
done
