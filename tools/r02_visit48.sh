# round 2, GPU visit 48: compact tile shapes
T=r02v48
O=gpurun_out
B="python bench.py --no-cpu-baseline --no-sweeps"
for tx in 79 80 81 82 83 84; do
$B --steps 8 --tile-x $tx > $O/${T}_128_tx${tx}.json 2>> $O/${T}_err.log; python -c "import json;d=json.load(open('$O/${T}_128_tx${tx}.json'));print('128 shape',$tx,'%.4e'%d['value'],'%.3f ms'%d['trace_ms_per_step'])"
done
for tx in 79 80 83; do
$B --n 256 --steps 2 --warmup 1 --tile-x $tx > $O/${T}_256_tx${tx}.json 2>> $O/${T}_err.log; python -c "import json;d=json.load(open('$O/${T}_256_tx${tx}.json'));print('256 shape',$tx,'%.4e'%d['value'],'%.3f ms'%d['trace_ms_per_step'])"
done
tail -3 $O/${T}_err.log
