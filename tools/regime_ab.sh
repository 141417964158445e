#!/bin/bash
mkdir -p gpurun_out
for v in "--warmup 3 --steps 50" "--warmup 100 --steps 50" "--warmup 200 --steps 50" "--warmup 400 --steps 50" "--warmup 800 --steps 50"; do
  timeout 300 python bench.py $v --no-cpu-baseline --no-realtime --headline-only --no-overlap > gpurun_out/rg_ab.json 2> gpurun_out/rg_ab.err
  python - "$v" <<'PY'
import json, sys
j = json.loads(open("gpurun_out/rg_ab.json").read().strip().splitlines()[-1])
r, e = j["roofline"], j["e2e"]
print(f"{sys.argv[1]:28s} value {j['value']/1e6:.3f} M ({j['ms_per_step']:.4f} ms, aec {r['kernel_ms_per_launch']:.4f} frac {r['frac']:.3f})  e2e {e['value']/1e6:.3f} M ({e['ms_per_step']:.4f} ms, aec {e.get('aec_ms_per_launch') or 0:.4f})")
PY
done
