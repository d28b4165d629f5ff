# round 2, GPU visit 22: A/B sort key with the direction octant
T=r02v22
O=gpurun_out
B="python bench.py --no-cpu-baseline --no-sweeps --steps 5"
for tx in 0 77; do
$B --tile-x $tx > $O/${T}_tx${tx}.json 2>> $O/${T}_err.log; python -c "import json;d=json.load(open('$O/${T}_tx${tx}.json'));print('128 sortMode',$tx,d['value'],d['trace_ms_per_step'])"
$B --tile-x $tx --n 256 --steps 2 > $O/${T}_256_tx${tx}.json 2>> $O/${T}_err.log; python -c "import json;d=json.load(open('$O/${T}_256_tx${tx}.json'));print('256 sortMode',$tx,d['value'],d['trace_ms_per_step'])"
done
tail -3 $O/${T}_err.log
