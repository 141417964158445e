#!/bin/bash
# A/B of the pipelined host path on one box. usage: tools/e2e_ab.sh "<env assignments>" ...
mkdir -p gpurun_out
for v in "$@"; do
  env $v timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-realtime --headline-only > gpurun_out/e2e_ab.json 2> gpurun_out/e2e_ab.err
  python - "$v" <<'PY'
import json, sys
try:
    j = json.loads(open("gpurun_out/e2e_ab.json").read().strip().splitlines()[-1])
    r, e, o = j["roofline"], j["e2e"], j.get("overlap_mode") or {}
    print(f"{sys.argv[1]:32s} value {j['value']/1e6:.3f} M ({j['ms_per_step']:.4f} ms, aec {r['kernel_ms_per_launch']:.4f} frac {r['frac']:.3f})  overlap {o.get('value', 0)/1e6:.3f} M  e2e {e['value']/1e6:.3f} M ({e['ms_per_step']:.4f} ms, aec {e.get('aec_ms_per_launch') or 0:.4f})")
except Exception as ex:
    print(sys.argv[1], "FAILED", ex, open("gpurun_out/e2e_ab.err").read()[-800:])
PY
done
