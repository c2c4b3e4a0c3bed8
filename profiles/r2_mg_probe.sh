# kernel times of the peer-memory operators at world = 1 (R-MAT scale $1, default 23)
S=${1:-23}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_mg1_launches.csv python profiles/mg_single_probe.py --scale $S > gpurun_out/r2_mg1.log 2>&1
tail -3 gpurun_out/r2_mg1.log
python profiles/summarize_launches.py gpurun_out/r2_mg1_launches.csv > gpurun_out/r2_mg1_launches.md 2>&1
head -30 gpurun_out/r2_mg1_launches.md
