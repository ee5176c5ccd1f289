set -x; mkdir -p gpurun_out
PE_CFG_ONLY_A=1 timeout 600 python tools/bench_configs.py > gpurun_out/r02_config_A.jsonl 2> gpurun_out/cfgA.err; cut -c1-330 gpurun_out/r02_config_A.jsonl; tail -2 gpurun_out/cfgA.err
timeout 900 python -m pytest tests/test_config_a.py -m gpu -q > gpurun_out/cfgA_tests.log 2>&1; tail -4 gpurun_out/cfgA_tests.log
