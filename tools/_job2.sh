# 2-GPU scratch job: data-parallel bench with the BatchNorm statistics exchange (K15 over NVLink; graphs on / off; collective fallback)
mkdir -p gpurun_out
run() { name=$1; shift; timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; tail -c 400 gpurun_out/$name.json | tr '\n' ' '; echo; }
run bench_2gpu_syncbn --sync-bn
run bench_2gpu_syncbn_eager --sync-bn --no-graphs
MAGGIE_B200_NO_PEER_EXCHANGE=1 run bench_2gpu_syncbn_nccl --sync-bn
