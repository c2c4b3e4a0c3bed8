# ncu --set full of ONE rcm_narrow_kernel launch (first BFS of a Poisson grid, cluster pinned to
# 16 so that the whole BFS is one launch) + per-line stall samples.  $1 = grid (default 2048)
G=${1:-2048}
TAG=${2:-r2a}
mkdir -p gpurun_out
SB200_RCM_CLUSTER=16 timeout 600 ncu --set full --clock-control none --import-source on -k regex:rcm_narrow_kernel -c 1 -f -o gpurun_out/${TAG}_rcm python profiles/prof_driver.py --ops rcm --graph poisson --grid $G --reps 1 > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
python profiles/ncu_summary.py gpurun_out/${TAG}_rcm.ncu-rep > gpurun_out/${TAG}_rcm_summary.md 2>/dev/null
ncu -i gpurun_out/${TAG}_rcm.ncu-rep --page source --print-source cuda,sass --csv -k regex:rcm_narrow > /tmp/rcm.csv 2>/dev/null
echo "## by stall samples" > gpurun_out/${TAG}_rcm_hotlines.md; python profiles/hotlines.py /tmp/rcm.csv 45 >> gpurun_out/${TAG}_rcm_hotlines.md 2>&1
echo "## by executed instructions" >> gpurun_out/${TAG}_rcm_hotlines.md; python profiles/hotlines.py /tmp/rcm.csv 30 inst >> gpurun_out/${TAG}_rcm_hotlines.md 2>&1
rm -f gpurun_out/${TAG}_rcm.ncu-rep
