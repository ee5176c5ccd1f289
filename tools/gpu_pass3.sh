set -x; mkdir -p gpurun_out
timeout 1300 python -m pytest tests -m gpu -q > gpurun_out/r02_gpu_tests.log 2>&1; tail -5 gpurun_out/r02_gpu_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; python tools/gpu_bench_line.py gpurun_out/r02_bench.json final; tail -2 gpurun_out/r02_bench.err
timeout 600 python tools/bench_configs.py > gpurun_out/r02_configs_C_D.jsonl 2> gpurun_out/cfgCD.err; cut -c1-330 gpurun_out/r02_configs_C_D.jsonl; tail -2 gpurun_out/cfgCD.err
./tools/micro/cusolver_point 1000 1024 > gpurun_out/r02_cusolver_point.json 2> gpurun_out/cusolver.err; cat gpurun_out/r02_cusolver_point.json; tail -2 gpurun_out/cusolver.err
