mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/c20_pytest.log 2>&1
tail -15 gpurun_out/c20_pytest.log
timeout 900 python profiles/bench_configs.py --config C4 --reps 3 --out gpurun_out/e_config_C4.json > gpurun_out/e_config_C4.log 2>&1; tail -c 1600 gpurun_out/e_config_C4.log
python profiles/tune_ops.py --graph poisson --size 4096 --ops coo_to_csr,csr_to_csc 2>&1 | tail -1
