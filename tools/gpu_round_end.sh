# what the driver runs at round end, plus the ncu evidence: tests, smoke, reference arm, bench, launch list, full capture
set -x; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/end_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/end_smoke.log 2>&1
python bench.py --impl reference > gpurun_out/end_ref.json 2> gpurun_out/end_ref.err
python bench.py > gpurun_out/end_bench.json 2> gpurun_out/end_bench.err &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/end_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/end_ncu_launches.log 2>&1
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --time-steps 20 > gpurun_out/end_plain20.json 2> gpurun_out/end_plain20.err &&
ncu --set full --clock-control none --import-source on -k regex:pe_b200 -s 1 -c 1 -o gpurun_out/end_prof python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --time-steps 20 > gpurun_out/end_ncu_full.log 2>&1
tail -2 gpurun_out/end_tests.log; tail -1 gpurun_out/end_smoke.log; head -c 300 gpurun_out/end_bench.json
