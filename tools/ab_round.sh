#!/bin/bash
# One A/B round on the GPU box (development aid): every build/variants/*.so over tools/ab_configs.py
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
{
for f in build/variants/*.so; do
  echo "== $(basename $f .so)"; DRTB_LIB=$PWD/$f python tools/ab_configs.py 2>&1 | tail -16
done
} > gpurun_out/ab_round.log 2>&1
tail -120 gpurun_out/ab_round.log
