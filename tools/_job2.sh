# 2-GPU scratch job: data-parallel bench, local BatchNorm statistics and the K15 statistics exchange
mkdir -p gpurun_out
run() { name=$1; shift; timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; grep "timed region" gpurun_out/$name.err; }
run bench_2gpu
run bench_2gpu_syncbn --sync-bn
