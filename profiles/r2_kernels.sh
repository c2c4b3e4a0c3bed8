# ncu --set full of every kernel of libsb200.so (one launch list + one full capture)
mkdir -p gpurun_out
timeout 1500 ncu --profile-from-start off --set full --clock-control none -f -o gpurun_out/r2_kernels python profiles/r2_kernels_driver.py > gpurun_out/r2_kernels.log 2>&1
tail -2 gpurun_out/r2_kernels.log
python profiles/ncu_summary.py gpurun_out/r2_kernels.ncu-rep > gpurun_out/r2_kernels_full.md 2>/dev/null
wc -l gpurun_out/r2_kernels_full.md
ls -la gpurun_out/r2_kernels.ncu-rep
