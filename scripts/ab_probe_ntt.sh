#!/bin/bash
# A/B probe of NTT library variants on one GPU: scripts/ab_probe_ntt.sh "<log sizes>" variant [variant ...]  ("base" = regular build)
cd "$(dirname "$0")/.."
SIZES=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = base ]; then unset BLAZE_B200_LIB; else export BLAZE_B200_LIB=$PWD/variants/$v.so; fi
  echo "== $v" | tee -a gpurun_out/ab_probe_ntt.log
  timeout 600 python scripts/ntt_probe.py $SIZES 2>&1 | grep log_n | cut -c1-300 | tee -a gpurun_out/ab_probe_ntt.log
done
