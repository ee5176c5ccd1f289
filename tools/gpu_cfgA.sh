set -x; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_config_a.py -m gpu -q > gpurun_out/cfgA_tests.log 2>&1; tail -4 gpurun_out/cfgA_tests.log
PE_CFG_ONLY_A=1 timeout 600 python tools/bench_configs.py > gpurun_out/r02_config_A.jsonl 2> gpurun_out/cfgA.err; cut -c1-420 gpurun_out/r02_config_A.jsonl; tail -2 gpurun_out/cfgA.err
PE_B200_FRONTAL_UPDATE_V1=1 PE_B200_FRONTAL_ONE_CTA_SUBST=1 PE_CFG_ONLY_A=1 timeout 600 python tools/bench_configs.py > gpurun_out/r02_config_A_v1.jsonl 2> gpurun_out/cfgA1.err; cut -c1-420 gpurun_out/r02_config_A_v1.jsonl | head -1
