set -x; mkdir -p gpurun_out
PE_CFG_ONLY_A=1 timeout 600 python tools/bench_configs.py > gpurun_out/r02_config_A.jsonl 2> gpurun_out/cfgA.err; cut -c1-330 gpurun_out/r02_config_A.jsonl; tail -2 gpurun_out/cfgA.err
PE_B200_FRONTAL_UPDATE_NARROW=1 PE_CFG_ONLY_A=1 timeout 600 python tools/bench_configs.py > gpurun_out/r02_config_A_narrow.jsonl 2> gpurun_out/cfgA1.err; cut -c1-330 gpurun_out/r02_config_A_narrow.jsonl | head -1
