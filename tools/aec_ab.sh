#!/bin/bash
# A/B of the AEC kernel builds on ONE box in ONE session: the headline step (4096 streams, 200 ticks) per variant,
# CUDA-event time of the AEC launches. usage: tools/aec_ab.sh <tag> "<env assignments>" ...
tag=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  name=$(echo "$v" | tr ' =' '__')
  env $v timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-realtime --headline-only \
    > gpurun_out/${tag}_ab_${name}.json 2> gpurun_out/${tag}_ab_${name}.err
  python - "$v" gpurun_out/${tag}_ab_${name}.json <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    r = j["roofline"]
    print(f"{sys.argv[1]:40s} aec {r['kernel_ms_per_launch']:.4f} ms/launch  frac {r['frac']:.3f}  step {j['ms_per_step']:.4f} ms  e2e {j['e2e']['ms_per_step']:.4f} ms")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
