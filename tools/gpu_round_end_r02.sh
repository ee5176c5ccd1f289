set -x; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_gpu_tests.log 2>&1; tail -5 gpurun_out/r02_gpu_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; python tools/gpu_bench_line.py gpurun_out/r02_bench.json final; tail -2 gpurun_out/r02_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_reference_arm.json 2> gpurun_out/r02_ref.err; cut -c1-200 gpurun_out/r02_reference_arm.json
timeout 600 python tools/bench_configs.py > gpurun_out/r02_configs_C_D.jsonl 2> gpurun_out/cfgCD.err; cut -c1-300 gpurun_out/r02_configs_C_D.jsonl
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
