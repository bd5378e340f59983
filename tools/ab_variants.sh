#!/bin/bash
# A/B the variant builds under build/variants (development aid)
for f in build/variants/*.so; do
  echo "== $f"
  DRTB_LIB=$PWD/$f python tools/ncu_target.py --precision f64 --reps 4 2>&1 | tail -1
  DRTB_LIB=$PWD/$f python tools/ncu_target.py --precision f32 --reps 4 2>&1 | tail -1
done
