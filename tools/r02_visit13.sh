# round 2, GPU visit 13: divq_sorted_kernel -- resident blocks per SM sweep, 32-KB carve-out
T=r02v13
O=gpurun_out
python -m pytest tests/test_gpu_indep.py -q -x -k "sorted or blocked" > $O/${T}_indep.log 2>&1; echo "indep rc=$?"; tail -3 $O/${T}_indep.log
B="python bench.py --no-cpu-baseline --no-sweeps --steps 5"
for k in 8 7 6 5 4; do
$B --blocked-march 3 --max-blocks-per-sm $k > $O/${T}_k${k}.json 2>> $O/${T}_err.log; python -c "import json;d=json.load(open('$O/${T}_k${k}.json'));print('128 bm 3 blocks/SM',$k,d['value'],d['trace_ms_per_step'])"
done
for k in 8 6 4; do
$B --n 256 --steps 2 --blocked-march 3 --max-blocks-per-sm $k > $O/${T}_256_k${k}.json 2>> $O/${T}_err.log; python -c "import json;d=json.load(open('$O/${T}_256_k${k}.json'));print('256 bm 3 blocks/SM',$k,d['value'],d['trace_ms_per_step'])"
done
tail -3 $O/${T}_err.log
