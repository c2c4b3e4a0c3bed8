set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/c6_pytest.log 2>&1
tail -3 gpurun_out/c6_pytest.log
python profiles/tune_ops.py --graph poisson --size 4096 --ops csr_to_csc,coo_sort,csr_to_coo >> gpurun_out/c6_tune.log 2>&1
python profiles/tune_ops.py --graph er --size 22 --ops csr_to_csc,coo_sort >> gpurun_out/c6_tune.log 2>&1
python profiles/tune_ops.py --graph rmat --size 23 --ops csr_to_csc,permute2d_deg,permute2d_rand,csr_to_coo,degree_reorder >> gpurun_out/c6_tune.log 2>&1
cat gpurun_out/c6_tune.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rs_downsweep -c 3 -f -o gpurun_out/c6_rs python profiles/prof_driver.py --ops csr_to_csc --graph er --grid 22 --reps 1 > gpurun_out/c6_ncu_rs.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c6_rmat_launches.csv python profiles/tune_ops.py --graph rmat --size 23 --ops permute2d_deg --reps 1 > gpurun_out/c6_rmat.log 2>&1
