# Permute2D under a random-like permutation (DegreeReorder of an R-MAT graph, the C4 shape):
# ncu --set full of the row-sort kernels (ss_tile_kernel, ss_big_kernel x2) with per-line stall
# samples.  $1 = R-MAT scale (default 24), $2 = tag
S=${1:-24}
TAG=${2:-r2_p2d}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ss_tile_kernel|ss_big_kernel' -c 3 -f -o gpurun_out/${TAG}_full python profiles/prof_driver.py --ops permute2d --graph rmat --grid $S --perm degree --reps 1 > gpurun_out/${TAG}_f.log 2>&1
tail -1 gpurun_out/${TAG}_f.log
python profiles/ncu_summary.py gpurun_out/${TAG}_full.ncu-rep > gpurun_out/${TAG}_full.md 2>/dev/null
cat gpurun_out/${TAG}_full.md
for K in ss_tile ss_big; do
  ncu -i gpurun_out/${TAG}_full.ncu-rep --page source --print-source cuda,sass --csv -k regex:$K > /tmp/p2d_$K.csv 2>/dev/null
  echo "## ${K}_kernel by stall samples" > gpurun_out/${TAG}_hotlines_$K.md; python profiles/hotlines.py /tmp/p2d_$K.csv 45 >> gpurun_out/${TAG}_hotlines_$K.md 2>&1
done
rm -f gpurun_out/${TAG}_full.ncu-rep
