set -x
mkdir -p gpurun_out
SB200_SHARD_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 3 > gpurun_out/m2_bench2.json 2> gpurun_out/m2_bench2.err
grep shard-trace gpurun_out/m2_bench2.json
