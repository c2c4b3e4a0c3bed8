# ncu --set full of ONE rcm_narrow_kernel launch (first BFS of C2) + per-line stall samples
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:rcm_narrow_kernel -c 1 -f -o gpurun_out/g_rcm python profiles/prof_driver.py --ops rcm --graph poisson --grid 4096 --reps 1 > gpurun_out/g_ncu.log 2>&1
tail -2 gpurun_out/g_ncu.log
python profiles/ncu_summary.py gpurun_out/g_rcm.ncu-rep > gpurun_out/g_rcm_summary.md 2>/dev/null
ncu -i gpurun_out/g_rcm.ncu-rep --page source --print-source cuda,sass --csv -k regex:rcm_narrow > /tmp/rcm.csv 2>/dev/null
echo "## by stall samples" > gpurun_out/g_rcm_hotlines.md; python profiles/hotlines.py /tmp/rcm.csv 40 >> gpurun_out/g_rcm_hotlines.md 2>&1
echo "## by executed instructions" >> gpurun_out/g_rcm_hotlines.md; python profiles/hotlines.py /tmp/rcm.csv 25 inst >> gpurun_out/g_rcm_hotlines.md 2>&1
ncu -i gpurun_out/g_rcm.ncu-rep --page details --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
h=rows[0]
for r in rows[1:]:
    sec,name,unit,val=r[h.index('Section Name')],r[h.index('Metric Name')],r[h.index('Metric Unit')],r[h.index('Metric Value')]
    if sec in ('GPU Speed Of Light Throughput','Scheduler Statistics','Compute Workload Analysis','Memory Workload Analysis','Occupancy','Warp State Statistics','Instruction Statistics','Launch Statistics'):
        print(sec[:22],'|',name,'|',unit,'|',val)
" > gpurun_out/g_rcm_details.txt
rm -f gpurun_out/g_rcm.ncu-rep
