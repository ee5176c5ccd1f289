set -x; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/v3_tests.log
for cfg in "128,2,2" "256,2,2" "256,2,1" "256,1,1" "128,1,1" "64,2,2" "512,1,1"; do
  timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --resident $cfg > gpurun_out/v3_bench_${cfg//,/_}.json 2> gpurun_out/v3_bench_${cfg//,/_}.err
done
tail -3 gpurun_out/v3_tests.log
for f in gpurun_out/v3_bench_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['resident'])"; done
