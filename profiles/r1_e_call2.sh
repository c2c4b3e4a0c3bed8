mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --kernel-name-base demangled -k regex:"sb200::" -c 120 -f -o gpurun_out/f_ops python profiles/prof_driver.py --ops rcm,permute2d,coo_sort,compressed_sort,features --graph rmat --grid 21 --reps 1 > gpurun_out/f_ncu_ops.log 2>&1
tail -2 gpurun_out/f_ncu_ops.log
python profiles/ncu_summary.py gpurun_out/f_ops.ncu-rep > gpurun_out/f_ncu_full_ops_rmat.md 2> gpurun_out/f_err.log
wc -l gpurun_out/f_ncu_full_ops_rmat.md
rm -f gpurun_out/f_ops.ncu-rep
