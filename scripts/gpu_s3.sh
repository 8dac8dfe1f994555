#!/bin/bash
# ad-hoc GPU session: tests, A/B of the fused two-product reduction, merged-table probe, bench, DFMA microbench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > $O/s3_tests.log 2>&1; tail -4 $O/s3_tests.log
export PROBE_CHECK=1
BLAZE_B200_LIB=$PWD/blaze_b200/libblaze_b200_nofuse.so BZ_MSM_PRECOMP=0 timeout 300 python scripts/perf_probe.py 24 0 > $O/s3_probe_nofuse.log 2>&1; tail -1 $O/s3_probe_nofuse.log
BZ_MSM_PRECOMP=0 timeout 300 python scripts/perf_probe.py 24 0 > $O/s3_probe_fuse.log 2>&1; tail -1 $O/s3_probe_fuse.log
BZ_MSM_PRECOMP=2 timeout 600 python scripts/perf_probe.py 24 0,20,22,23,24 > $O/s3_probe_merged.log 2>&1; tail -5 $O/s3_probe_merged.log
(time timeout 900 python bench.py) > $O/s3_bench.log 2>&1; tail -2 $O/s3_bench.log | cut -c1-1500
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/imad_mb scripts/imad_microbench.cu && timeout 120 /tmp/imad_mb > $O/s3_microbench.txt 2>&1; grep -i "dfma" $O/s3_microbench.txt
nvidia-smi --query-gpu=memory.used,memory.total --format=csv,noheader
