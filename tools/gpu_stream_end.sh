set -x; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_stream.py -m gpu -q > gpurun_out/pass9_tests.log 2>&1; tail -3 gpurun_out/pass9_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; python tools/gpu_bench_line.py gpurun_out/r02_bench.json final; tail -2 gpurun_out/r02_bench.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
