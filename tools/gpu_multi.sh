set -x; mkdir -p gpurun_out
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/multi_${N}.json 2> gpurun_out/multi_${N}.err
tail -c 600 gpurun_out/multi_${N}.json
python bench.py --steps 3 --warmup 3 > gpurun_out/multi_1.json 2> gpurun_out/multi_1.err
