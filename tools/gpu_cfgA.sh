set -x; mkdir -p gpurun_out
PE_CFG_ONLY_A=1 PE_CFG_A_NO_REF=1 timeout 600 python tools/bench_configs.py > gpurun_out/r02_config_A_la.jsonl 2> gpurun_out/cfgA.err; cut -c1-330 gpurun_out/r02_config_A_la.jsonl; tail -2 gpurun_out/cfgA.err
PE_B200_FRONTAL_NO_LOOKAHEAD=1 PE_CFG_ONLY_A=1 PE_CFG_A_NO_REF=1 timeout 600 python tools/bench_configs.py > gpurun_out/r02_config_A_nola.jsonl 2> gpurun_out/cfgA1.err; cut -c1-330 gpurun_out/r02_config_A_nola.jsonl | head -1
timeout 900 python -m pytest tests/test_config_a.py -m gpu -q > gpurun_out/cfgA_tests.log 2>&1; tail -4 gpurun_out/cfgA_tests.log
