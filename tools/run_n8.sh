#!/bin/bash
# The 8-GPU records of a round (run with `gpurun --gpus 8`): scaling bench, config 3 with both gradient exchanges.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
N=${1:-8}
$TR --nproc-per-node $N --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_scale_n${N}_b7.json 2> gpurun_out/r02_scale_n${N}_b7.err
cut -c1-300 gpurun_out/r02_scale_n${N}_b7.json
{
for size in 256 512; do
  echo "== inverse loop ${size}^2, 64 spp, 4 bounces, 100 iterations"
  python examples/inverse_render.py --size $size --quiet 2>&1 | tail -1
  $TR --nproc-per-node $N --master-port 29512 examples/inverse_render.py --size $size --quiet 2>&1 | tail -1
  $TR --nproc-per-node $N --master-port 29513 examples/inverse_render.py --size $size --quiet --nccl 2>&1 | tail -1
done
} > gpurun_out/r02_inverse_render_n${N}.log 2>&1
cat gpurun_out/r02_inverse_render_n${N}.log
