# ncu --set full + per-line stalls of mg_push_rows_kernel at world = 1 (all stores local)
S=${1:-22}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mg_push_rows -c 1 -f -o gpurun_out/r2_push python profiles/mg_single_probe.py --scale $S > gpurun_out/r2_push.log 2>&1
tail -2 gpurun_out/r2_push.log
python profiles/ncu_summary.py gpurun_out/r2_push.ncu-rep > gpurun_out/r2_push_summary.md 2>/dev/null
cat gpurun_out/r2_push_summary.md
ncu -i gpurun_out/r2_push.ncu-rep --page source --print-source cuda,sass --csv -k regex:mg_push > /tmp/push.csv 2>/dev/null
python profiles/hotlines.py /tmp/push.csv 25 > gpurun_out/r2_push_hotlines.md 2>&1
cat gpurun_out/r2_push_hotlines.md
ncu -i gpurun_out/r2_push.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
h=rows[0]
want=['dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sectors_op_write.sum','lts__t_sectors_op_read.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum','l1tex__t_requests_pipe_lsu_mem_global_op_st.sum','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','gpu__time_duration.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum']
for r in rows[2:]:
    for w in want:
        if w in h: print(w, r[h.index(w)])
"
rm -f gpurun_out/r2_push.ncu-rep
