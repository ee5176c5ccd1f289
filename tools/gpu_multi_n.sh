# usage: bash tools/gpu_multi_n.sh N   -- config D sharded over N GPUs (timed NCCL all-gather) + the bench line at N GPUs
set -x; mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tools/config_d_multi.py > gpurun_out/r02_config_D_${N}gpu.json 2> gpurun_out/cfgD${N}.err
cut -c1-900 gpurun_out/r02_config_D_${N}gpu.json; tail -2 gpurun_out/cfgD${N}.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/b${N}.err
python tools/gpu_bench_line.py gpurun_out/r02_bench_${N}gpu.json ${N}gpu; tail -2 gpurun_out/b${N}.err
