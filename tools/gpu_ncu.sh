# usage: bash tools/gpu_ncu.sh <tag> <bench args...>   -- plain run first, then launch list + one full-set capture
set -x; mkdir -p gpurun_out
tag=$1; shift
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e "$@" > gpurun_out/${tag}_plain.json 2> gpurun_out/${tag}_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:pe_b200 -s 1 -c 1 -o gpurun_out/${tag}_prof python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e "$@" > gpurun_out/${tag}_ncu.log 2>&1
cat gpurun_out/${tag}_plain.json | head -c 400
