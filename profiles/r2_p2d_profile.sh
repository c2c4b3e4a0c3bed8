# Permute2D under a random-like permutation (DegreeReorder of an R-MAT graph, the C4 shape):
# ncu --set full of the gather kernel (ss_tile_kernel) and the long-row fill / store kernels,
# per-line stall samples of ss_tile_kernel.  $1 = R-MAT scale (default 24)
S=${1:-24}
TAG=${2:-r2_p2d}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ss_tile_kernel|ss_long_fill|ss_long_store' -c 3 -f -o gpurun_out/${TAG}_full python profiles/prof_driver.py --ops permute2d --graph rmat --grid $S --perm degree --reps 1 > gpurun_out/${TAG}_f.log 2>&1
tail -1 gpurun_out/${TAG}_f.log
python profiles/ncu_summary.py gpurun_out/${TAG}_full.ncu-rep > gpurun_out/${TAG}_full.md 2>/dev/null
cat gpurun_out/${TAG}_full.md
ncu -i gpurun_out/${TAG}_full.ncu-rep --page source --print-source cuda,sass --csv -k regex:ss_tile > /tmp/p2d.csv 2>/dev/null
echo "## ss_tile_kernel by stall samples" > gpurun_out/${TAG}_hotlines.md; python profiles/hotlines.py /tmp/p2d.csv 40 >> gpurun_out/${TAG}_hotlines.md 2>&1
rm -f gpurun_out/${TAG}_full.ncu-rep
