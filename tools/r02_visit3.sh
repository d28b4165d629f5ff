# round 2, GPU visit 3: the blocked independent-ray kernel -- parity against the one-cell kernel, A/B throughput, ncu
T=r02v3
O=gpurun_out
python -m pytest tests/test_gpu_indep.py -q -x > $O/${T}_indep.log 2>&1; echo "indep rc=$?"; tail -12 $O/${T}_indep.log
python -m pytest tests/test_gpu_atsize.py -q -s > $O/${T}_atsize.log 2>&1; echo "atsize rc=$?"; grep -E "passed|failed" $O/${T}_atsize.log | tail -2
B="python bench.py --no-cpu-baseline --no-sweeps --steps 5"
for rl in 0 4 16; do
$B --refill-lanes $rl        > $O/${T}_blk_rl${rl}.json 2>> $O/${T}_err.log; python -c "import json;d=json.load(open('$O/${T}_blk_rl${rl}.json'));print('refill',$rl,d['value'],d['ms_per_step'],d['trace_ms_per_step'])"
done
$B --blocked-march 0         > $O/${T}_v1.json 2>> $O/${T}_err.log; python -c "import json;d=json.load(open('$O/${T}_v1.json'));print('v1',d['value'],d['ms_per_step'])"
$B --n 256 --steps 2         > $O/${T}_blk_256.json 2>> $O/${T}_err.log; python -c "import json;d=json.load(open('$O/${T}_blk_256.json'));print('256',d['value'],d['ms_per_step'])"
ncu --set full --clock-control none --import-source on -k regex:divq_indep_kernel -s 3 -c 1 -f -o $O/${T}_divq_indep_128 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sweeps > $O/${T}_ncu1.log 2>&1
tail -3 $O/${T}_err.log
