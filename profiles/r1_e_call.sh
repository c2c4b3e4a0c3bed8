mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/c27_pytest.log 2>&1
tail -5 gpurun_out/c27_pytest.log
python profiles/tune_ops.py --graph rmat --size 23 --ops permute2d_deg,permute2d_rand,csr_to_csc 2>&1 | tail -1
python profiles/tune_ops.py --graph rmat --size 25 --ops permute2d_deg 2>&1 | tail -1
python profiles/tune_ops.py --graph poisson --size 4096 --ops csr_to_csc,coo_sort 2>&1 | tail -1
