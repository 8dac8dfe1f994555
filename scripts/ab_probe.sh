#!/bin/bash
# A/B probe of library variants on one GPU: scripts/ab_probe.sh "<sizes>" variant [variant ...]   ("base" = the regular build)
cd "$(dirname "$0")/.."
SIZES=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = base ]; then unset BLAZE_B200_LIB; else export BLAZE_B200_LIB=$PWD/variants/$v.so; fi
  echo "== $v" | tee -a gpurun_out/ab_probe.log
  timeout 300 python scripts/perf_probe.py $SIZES 0 2>&1 | grep logn | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); p = r['phase_ms']
    print(r['logn'], 'ok', r['ok'], 'c', r['plan']['c'], 'merged', r['plan']['merged_table'], 'total %.2f sort %.2f acc %.2f reduce %.2f' % (p['total'], p['sort'], p['accumulate'], p['reduce']))
" | tee -a gpurun_out/ab_probe.log
done
