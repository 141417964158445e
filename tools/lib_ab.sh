#!/bin/bash
# A/B of whole library builds on ONE box in ONE session: the headline step per build (MSB200_LIB selects the .so),
# interleaved and repeated. usage: tools/lib_ab.sh <tag> <reps> name=path.so ...
tag=$1; reps=$2; shift 2
mkdir -p gpurun_out
for r in $(seq 1 $reps); do
for v in "$@"; do
  name=${v%%=*}; lib=${v#*=}
  MSB200_LIB=$lib timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-realtime --headline-only \
    > gpurun_out/${tag}_ab_${name}_$r.json 2> gpurun_out/${tag}_ab_${name}_$r.err
  python - "$name" gpurun_out/${tag}_ab_${name}_$r.json <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    r = j["roofline"]
    su = j.get("startup_regime", {})
    print(f"{sys.argv[1]:14s} aec {r['kernel_ms_per_launch']:.4f} ms/launch  frac {r['frac']:.3f}  step {j['ms_per_step']:.4f} ms  e2e {j['e2e']['ms_per_step']:.4f} ms  startup aec {su.get('aec_ms_per_launch', float('nan')):.4f}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
done
