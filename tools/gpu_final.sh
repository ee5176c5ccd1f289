set -x; mkdir -p gpurun_out
nproc > gpurun_out/final_nproc.txt
( time python bench.py ) > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
tail -2 gpurun_out/final_smoke.log
