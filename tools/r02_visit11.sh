# round 2, GPU visit 11: blockedMarch = 3 (divq_sorted_kernel: length-sorted independent rays on the blocked copy) -- parity, A/B, counters
T=r02v11
O=gpurun_out
python -m pytest tests/test_gpu_indep.py -q -x > $O/${T}_indep.log 2>&1; echo "indep rc=$?"; tail -5 $O/${T}_indep.log
B="python bench.py --no-cpu-baseline --no-sweeps --steps 5"
for bm in 2 3; do
$B --blocked-march $bm > $O/${T}_bm${bm}.json 2>> $O/${T}_err.log; python -c "import json;d=json.load(open('$O/${T}_bm${bm}.json'));print('128 bm',$bm,d['value'],d['ms_per_step'],d['trace_ms_per_step'])"
done
for rp in 25 50; do
$B --blocked-march 3 --refill-lanes $rp > $O/${T}_bm3_rp${rp}.json 2>> $O/${T}_err.log; python -c "import json;d=json.load(open('$O/${T}_bm3_rp${rp}.json'));print('128 bm 3 rp',$rp,d['value'],d['ms_per_step'],d['trace_ms_per_step'])"
done
$B --n 256 --steps 2 --blocked-march 3 > $O/${T}_256_bm3.json 2>> $O/${T}_err.log; python -c "import json;d=json.load(open('$O/${T}_256_bm3.json'));print('256 bm 3',d['value'],d['ms_per_step'],d['trace_ms_per_step'])"
M=l1tex__data_pipe_lsu_wavefronts.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sector_hit_rate.pct,l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum
ncu --metrics $M --clock-control none -k regex:divq_sorted_kernel -s 3 -c 1 --csv --log-file $O/${T}_sorted_counters.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sweeps --blocked-march 3 > /dev/null 2>&1
python - <<'PY'
import csv
for f in ("gpurun_out/r02v11_sorted_counters.csv",):
    try:
        rows=[r for r in csv.reader(open(f)) if len(r)>10 and r[0].isdigit()]
        print(f); [print("  ",r[-3],r[-1]) for r in rows]
    except Exception as e: print(f,e)
PY
tail -3 $O/${T}_err.log
