set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/c14_pytest.log 2>&1
tail -3 gpurun_out/c14_pytest.log
python profiles/block_probe.py
python profiles/tune_ops.py --graph poisson --size 4096 --ops coo_to_csr,csr_to_csc,permute2d_rcm
