# Round-1 (session e) evidence run on one B200: tests, bench (both arms), launch list, ncu --set full
# of the operator kernels, per-config tables.  Outputs under gpurun_out/e_*.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem --format=csv
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/e_pytest.log 2>&1
tail -3 gpurun_out/e_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/e_smoke.log 2>&1; tail -2 gpurun_out/e_smoke.log
timeout 900 python bench.py > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/e_bench_reference_arm.json 2>> gpurun_out/e_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:"sb200::" -c 400 --csv --log-file gpurun_out/e_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/e_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"sb200::(permute|scan_|rs_|boundary|expand|degree|ss_|ex_|gap|inverse|coo_)" -c 40 -f -o gpurun_out/e_ops python profiles/prof_driver.py --ops rcm,permute2d,csr_to_csc,coo_to_csr,degree_reorder --graph poisson --grid 4096 --reps 1 > gpurun_out/e_ncu_ops.log 2>&1
python profiles/ncu_summary.py gpurun_out/e_ops.ncu-rep > gpurun_out/e_ncu_full_ops.md 2> gpurun_out/e_ncu_summary.err
for k in permute_short_rows rs_downsweep_pipe boundary_fill_copy_vec expand_ptr permute_prepare degree_downsweep; do
  ncu -i gpurun_out/e_ops.ncu-rep --page source --print-source cuda,sass --csv -k regex:$k > /tmp/src_$k.csv 2>/dev/null
  echo "## $k" >> gpurun_out/e_hotlines.md; python profiles/hotlines.py /tmp/src_$k.csv 12 >> gpurun_out/e_hotlines.md 2>&1
done
rm -f gpurun_out/e_ops.ncu-rep
for c in C1 C3 C4; do timeout 900 python profiles/bench_configs.py --config $c --reps 3 --out gpurun_out/e_config_$c.json > gpurun_out/e_config_$c.log 2>&1; tail -c 300 gpurun_out/e_config_$c.log; done
ls -la gpurun_out | tail -20
