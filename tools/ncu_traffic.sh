#!/bin/bash
# Regenerates profiles/ncu_render_kernel.json (bench.py's roofline.traffic) for the headline workload:
#   on the GPU box   : bash tools/ncu_traffic.sh capture     -> gpurun_out/traffic_f64.ncu-rep
#   back in the repo : bash tools/ncu_traffic.sh summarize [name]
# The JSON entry records the hash of the kernel sources; bench.py refuses an entry taken at other sources.
set -e
case "$1" in
  capture)
    ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 2 -c 1 -f -o gpurun_out/traffic_f64 \
        python tools/ncu_target.py --precision f64 --reps 3 ;;
  summarize)
    python tools/summarize_ncu.py gpurun_out/traffic_f64.ncu-rep "${2:-r02_f64_bench_kernel}" f64_1024x1024_256spp_b8 ;;
  *) echo "usage: $0 capture|summarize [name]"; exit 1 ;;
esac
