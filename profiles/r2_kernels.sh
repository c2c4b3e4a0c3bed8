# ncu --set full of every kernel of libsb200.so (profiles/r2_kernels_driver.py), summarised per
# launch and per kernel.  The capture is cut at 450 launches (the wide RCM regime at the end
# repeats the same kernels per level).
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --set full --clock-control none -c 450 -f -o gpurun_out/r2_kernels python profiles/r2_kernels_driver.py > gpurun_out/r2_kernels.log 2>&1
tail -2 gpurun_out/r2_kernels.log
python profiles/ncu_summary.py gpurun_out/r2_kernels.ncu-rep > gpurun_out/r2_kernels_full.md 2>/dev/null
python profiles/ncu_by_kernel.py gpurun_out/r2_kernels_full.md > gpurun_out/r2_kernels_by_kernel.md
wc -l gpurun_out/r2_kernels_full.md gpurun_out/r2_kernels_by_kernel.md
rm -f gpurun_out/r2_kernels.ncu-rep
