#!/bin/bash
# One gpurun session: GPU parity tests, the bench line, the ncu launch list of the same command and ncu --set full
# captures of the two dominant kernels. Outputs land in gpurun_out/ (scratch); summaries are copied to profiles/.
# usage: tools/gpu_session.sh <tag> [what...]   what: tests bench launches aec scale (default: all)
tag=${1:-rX}; shift
what=${*:-tests bench launches aec scale}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
for w in $what; do
case $w in
tests) timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests exit $?" ; tail -3 gpurun_out/${tag}_tests.log ;;
bench) timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"; cat gpurun_out/${tag}_bench.json ;;
launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-realtime > gpurun_out/${tag}_launches.log 2>&1; echo "launches exit $?" ;;
aec) timeout 600 ncu --set full --clock-control none --import-source on -k regex:aec_kernel -s 4 -c 1 -f -o gpurun_out/${tag}_aec python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-realtime > gpurun_out/${tag}_aec.log 2>&1; echo "aec exit $?" ;;
scale) timeout 600 ncu --set full --clock-control none --import-source on -k regex:scale_rgb -s 2 -c 1 -f -o gpurun_out/${tag}_scale python bench_video.py --frames 512 --iters 3 > gpurun_out/${tag}_scale.log 2>&1; echo "scale exit $?" ;;
*) echo "running custom: $w"; timeout 600 bash -c "$w" ;;
esac
done
