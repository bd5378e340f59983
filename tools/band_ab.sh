#!/bin/bash
# A/B of the shard band size at N GPUs (development aid): tools/band_ab.sh N "1 2 4"
N=$1
for b in $2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 20 --warmup 3 --band-rows $b 2>/dev/null > /tmp/band_$b.json
  python - "$b" <<'PY'
import json, sys
d = json.load(open(f"/tmp/band_{sys.argv[1]}.json"))
print("band_rows", sys.argv[1], "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1))
PY
done
