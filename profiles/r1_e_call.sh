mkdir -p gpurun_out
python profiles/tune_ops.py --graph rmat --size 23 --ops permute2d_deg,permute2d_rand 2>&1 | tail -1
SB200_P2D_MID=0 python profiles/tune_ops.py --graph er --size 24 --ops permute2d_deg 2>&1 | tail -1
timeout 300 python -m pytest tests -m gpu -x -q -k "permute2d or compressed or hubs" 2>&1 | tail -2
