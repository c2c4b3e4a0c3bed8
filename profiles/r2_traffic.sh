# DRAM traffic of ONE Permute2D call on C4 (R-MAT-26, DegreeReorder permutation): the dram__bytes
# counters of every kernel of the call (the same counters --set full collects; the full sections of
# these kernels are in r2c_p2d_full.md at scale 24) -> profiles/r2_traffic.json (bench.py reports it as roofline.traffic)
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,launch__grid_size,launch__block_size --clock-control none -f -o gpurun_out/r2_traffic python profiles/p2d_time.py --scale 26 --reps 1 > gpurun_out/r2_traffic.log 2>&1
grep P2D_ gpurun_out/r2_traffic.log
python profiles/ncu_summary.py gpurun_out/r2_traffic.ncu-rep > gpurun_out/r2_traffic_full.md 2>/dev/null
python profiles/traffic_json.py gpurun_out/r2_traffic_full.md > gpurun_out/r2_traffic.json
cat gpurun_out/r2_traffic.json | head -12
rm -f gpurun_out/r2_traffic.ncu-rep
