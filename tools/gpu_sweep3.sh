# usage: bash tools/gpu_sweep3.sh [notests] "tag|ENV=.. |bench args" ...   (env part optional: tag|args or tag|env|args)
set -x; mkdir -p gpurun_out
if [ "$1" != "notests" ]; then
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/sweep_tests.log
else shift; fi
for cfg in "$@"; do
  tag="${cfg%%|*}"; rest="${cfg#*|}"
  if [[ "$rest" == *"|"* ]]; then envs="${rest%%|*}"; args="${rest#*|}"; else envs=""; args="$rest"; fi
  env $envs timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e $args > gpurun_out/sweep_${tag}.json 2> gpurun_out/sweep_${tag}.err
done
tail -3 gpurun_out/sweep_tests.log
