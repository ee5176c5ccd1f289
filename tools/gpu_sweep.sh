# usage: bash tools/gpu_sweep.sh [notests] cfg...   where cfg = S,I,J,workspace
set -x; mkdir -p gpurun_out
if [ "$1" != "notests" ]; then
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/sweep_tests.log
else shift; fi
for cfg in "$@"; do
  IFS=, read S I J W <<< "$cfg"
  timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --resident $S,$I,$J --workspace $W > gpurun_out/sweep_${S}_${I}_${J}_${W}.json 2> gpurun_out/sweep_${S}_${I}_${J}_${W}.err
done
tail -3 gpurun_out/sweep_tests.log
