#!/bin/bash
# throughput and median stage times at the kernel sizes of BASELINE config 5
for k in 11 31 71; do
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --k $k --t0 -3 --t1 13 2>/dev/null > /tmp/sweep_$k.json
  K=$k python - <<'PY'
import json, os
k = os.environ["K"]
d = json.loads(open(f"/tmp/sweep_{k}.json").read().strip().splitlines()[-1])
print("k", k, round(d["value"], 1), [(x["kernel"].split(".")[-1], x["ms"]) for x in d["roofline"]["kernels"][:4]])
PY
done
