#!/bin/bash
# A/B of the chain's overlap mode on one box: bench headline with and without it
mkdir -p gpurun_out
for v in "--no-overlap" "" "--no-overlap" ""; do
  MSB200_CHAIN_OVERLAP=$([ -z "$v" ] && echo 1 || echo 0) timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-realtime --headline-only $v > gpurun_out/ov_ab.json 2> gpurun_out/ov_ab.err
  python - "$v" <<'PY'
import json, sys
try:
    j = json.loads(open("gpurun_out/ov_ab.json").read().strip().splitlines()[-1])
    r = j["roofline"]
    print(f"{sys.argv[1] or 'overlap':14s} value {j['value']/1e6:.3f} M  step {j['ms_per_step']:.4f} ms  e2e {j['e2e']['value']/1e6:.3f} M ({j['e2e']['ms_per_step']:.4f} ms)  aec {r['kernel_ms_per_launch']:.4f} ms frac {r['frac']:.3f} share {r['kernel_share_of_step']:.3f}")
except Exception as e:
    print(sys.argv[1], "FAILED", e, open("gpurun_out/ov_ab.err").read()[-800:])
PY
done
