set -x
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_sharded_gpu.py -x -q -m gpu ) > gpurun_out/m3_pytest.log 2>&1
tail -3 gpurun_out/m3_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/m3_bench2.json 2> gpurun_out/m3_bench2.err
tail -3 gpurun_out/m3_bench2.err
