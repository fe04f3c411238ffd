run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu --no-e2e $3 2>&1 | tail -1 | python scripts/benchline.py "$1"; }
run peer-overlap 29511
run peer-no-overlap 29512 --no-overlap
run nccl-no-overlap 29513 "--no-overlap --transport nccl"
run peer-overlap 29514
