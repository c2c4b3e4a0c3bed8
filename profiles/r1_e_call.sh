mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/c39_pytest.log 2>&1
tail -4 gpurun_out/c39_pytest.log
python profiles/tune_ops.py --graph poisson --size 4096 --ops csr_to_csc,coo_sort,permute2d_rcm 2>&1 | tail -1
