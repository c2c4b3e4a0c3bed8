mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/c23_pytest.log 2>&1
tail -5 gpurun_out/c23_pytest.log
python profiles/tune_ops.py --graph poisson --size 4096 --ops csr_to_csc,coo_sort 2>&1 | tail -1
python profiles/tune_ops.py --graph er --size 24 --ops csr_to_csc,coo_sort,permute2d_deg 2>&1 | tail -1
python profiles/tune_ops.py --graph rmat --size 23 --ops csr_to_csc,coo_sort,permute2d_deg,rcm 2>&1 | tail -1
SB200_RS_CONFIG=5 SB200_RS_CHUNKS_PER_SM=4 python profiles/tune_ops.py --graph rmat --size 23 --ops csr_to_csc,coo_sort,permute2d_deg,rcm 2>&1 | tail -1
