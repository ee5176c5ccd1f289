# one-GPU round-end pass: GPU parity tests, bench line, reference arm, configs, ncu of the complex tree kernel (config D)
set -x; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests.log 2>&1; tail -3 gpurun_out/r02_gpu_tests.log
python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; python tools/gpu_bench_line.py gpurun_out/r02_bench.json final; tail -2 gpurun_out/r02_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_reference_arm.json 2> gpurun_out/r02_ref.err; cut -c1-300 gpurun_out/r02_reference_arm.json
PE_CFG_ONLY_A=1 timeout 900 python tools/bench_configs.py > gpurun_out/r02_config_A.jsonl 2> gpurun_out/cfgA.err; cut -c1-600 gpurun_out/r02_config_A.jsonl; tail -2 gpurun_out/cfgA.err
timeout 600 python tools/bench_configs.py > gpurun_out/r02_configs_C_D.jsonl 2> gpurun_out/cfgCD.err; cut -c1-400 gpurun_out/r02_configs_C_D.jsonl
PE_CFG_ONLY_AC=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:pe_b200_tree -s 1 -c 1 -o gpurun_out/r02_tree_cplx_d python tools/bench_configs.py > gpurun_out/ncu_d.log 2>&1; tail -3 gpurun_out/ncu_d.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
