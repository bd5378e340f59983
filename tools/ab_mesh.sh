#!/bin/bash
# A/B of build/variants/*.so on config 4 (development aid): tools/mesh_bench.py per variant, two rounds
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
{
for round in 1; do
for f in build/variants/*.so; do
  echo "== $(basename $f .so) round $round"
  DRTB_LIB=$PWD/$f python tools/mesh_bench.py --reps 3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('f64') or l.startswith('f32'):
        d = json.loads(l[4:]); print(l[:3], round(d['kernel_ms'],2), 'ms', round(d['Msegments_per_s'],1), 'Mseg/s nodes', round(d['bvh_nodes_per_segment'],2), 'tests', round(d['tri_tests_per_segment'],2), 'mean', d['image_mean'])
"
done
done
} > gpurun_out/ab_mesh.log 2>&1
cat gpurun_out/ab_mesh.log
