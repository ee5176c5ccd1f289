# usage: bash tools/gpu_ncu_ab.sh "tag|bench args" ...  : ncu --set full capture (20 time steps) of the solve kernel per variant
set -x; mkdir -p gpurun_out
for cfg in "$@"; do
  tag="${cfg%%|*}"; args="${cfg#*|}"
  timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --time-steps 20 $args > gpurun_out/ab_${tag}_plain.json 2> gpurun_out/ab_${tag}_plain.err &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:pe_b200 -s 1 -c 1 -f -o gpurun_out/ab_${tag} python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --time-steps 20 $args > gpurun_out/ab_${tag}_ncu.log 2>&1
done
