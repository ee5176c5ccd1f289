set -x; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pivot_guard.py tests/test_full_size.py tests/test_checkpoint.py tests/test_sharding.py -m gpu -q > gpurun_out/pass5_tests.log 2>&1; tail -4 gpurun_out/pass5_tests.log
timeout 900 python -m pytest tests/test_parity.py -m gpu -q -k "ac_" > gpurun_out/pass5_ac.log 2>&1; tail -3 gpurun_out/pass5_ac.log
timeout 600 python tools/bench_configs.py > gpurun_out/r02_configs_C_D.jsonl 2> gpurun_out/cfgCD.err; cut -c1-300 gpurun_out/r02_configs_C_D.jsonl
