mkdir -p gpurun_out
timeout 900 python profiles/bench_configs.py --config C5 --rcm --reps 2 --out gpurun_out/c37_C5.json > gpurun_out/c37_C5.log 2>&1; tail -c 300 gpurun_out/c37_C5.log
