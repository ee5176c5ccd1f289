# usage: bash tools/gpu_bench_n.sh N   -- the bench line at N GPUs of one box
set -x; mkdir -p gpurun_out
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/b${N}.err
python tools/gpu_bench_line.py gpurun_out/r02_bench_${N}gpu.json ${N}gpu; tail -2 gpurun_out/b${N}.err
