# round 2, GPU visit 23: RESORT (first legs sorted, parked rays sorted again) -- parity tests, C5 A/B, 128^3 unchanged
T=r02v23
O=gpurun_out
python -m pytest tests/test_gpu_indep.py tests/test_gpu_parity.py tests/test_gpu_atsize.py -m gpu -q -x > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $O/${T}_pytest.log
python tools/bench_configs.py --config C5 --steps 2 > $O/${T}_c5.json 2> $O/${T}_c5.err; cut -c1-420 $O/${T}_c5.json
python bench.py --no-cpu-baseline --no-sweeps --steps 5 > $O/${T}_128.json 2>> $O/${T}_err.log; python -c "import json;d=json.load(open('$O/${T}_128.json'));print('128',d['value'],d['trace_ms_per_step'])"
M=smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__t_sector_hit_rate.pct,l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum
ncu --metrics $M --clock-control none -k regex:divq_sorted_kernel -s 2 -c 2 --csv --log-file $O/${T}_c5_counters.csv python tools/bench_configs.py --config C5 --n 128 --steps 1 > /dev/null 2>&1
python - <<'PY'
import csv
for f in ("gpurun_out/r02v23_c5_counters.csv",):
    try:
        rows=[r for r in csv.reader(open(f)) if len(r)>10 and r[0].isdigit()]
        print(f); [print("  ",r[4][:60],r[-3],r[-1]) for r in rows]
    except Exception as e: print(f,e)
PY
tail -3 $O/${T}_c5.err
