#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
export PROBE_CHECK=1 BZ_MSM_PRECOMP=2
timeout 300 python scripts/perf_probe.py 25 0 > $O/s11_probe_a.log 2>&1; tail -1 $O/s11_probe_a.log | cut -c1-420
BLAZE_B200_LIB=$PWD/blaze_b200/libblaze_b200_pt4k.so timeout 300 python scripts/perf_probe.py 25 0 > $O/s11_probe_pt4k.log 2>&1; tail -1 $O/s11_probe_pt4k.log | cut -c1-420
timeout 300 python scripts/ntt_probe2.py 27 3 > $O/s11_ntt.log 2>&1; tail -2 $O/s11_ntt.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_ntt_pass -s 3 -c 1 -f -o $O/s11_ntt_pass_2p24 python scripts/ntt_probe2.py 24 2 > $O/s11_ncu_ntt.log 2>&1
tail -2 $O/s11_ncu_ntt.log
