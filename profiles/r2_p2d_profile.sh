# Permute2D under a random-like permutation (DegreeReorder of an R-MAT graph, the C4 shape):
# launch list of one call + ncu --set full of its kernels.  $1 = R-MAT scale (default 24)
S=${1:-24}
TAG=${2:-r2_p2d}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python profiles/prof_driver.py --ops permute2d --graph rmat --grid $S --perm degree --reps 1 > gpurun_out/${TAG}_l.log 2>&1
python profiles/summarize_launches.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches.md 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ss_tile|ss_long|rs_down|rs_up|permute_prepare|scan_lookback' -c 12 -f -o gpurun_out/${TAG}_full python profiles/prof_driver.py --ops permute2d --graph rmat --grid $S --perm degree --reps 1 > gpurun_out/${TAG}_f.log 2>&1
python profiles/ncu_summary.py gpurun_out/${TAG}_full.ncu-rep > gpurun_out/${TAG}_full.md 2>/dev/null
ncu -i gpurun_out/${TAG}_full.ncu-rep --page source --print-source cuda,sass --csv -k regex:ss_tile > /tmp/p2d.csv 2>/dev/null
echo "## ss_tile_kernel by stall samples" > gpurun_out/${TAG}_hotlines.md; python profiles/hotlines.py /tmp/p2d.csv 40 >> gpurun_out/${TAG}_hotlines.md 2>&1
rm -f gpurun_out/${TAG}_full.ncu-rep
