# round 2, GPU visit 47: A/B compact tiles (8 x 2 x 2 origin cells) against rows of 32 cells
T=r02v47
O=gpurun_out
B="python bench.py --no-cpu-baseline --no-sweeps"
for rep in 1 2; do for tx in 0 79; do
$B --steps 10 --tile-x $tx > $O/${T}_128_tx${tx}.json 2>> $O/${T}_err.log; python -c "import json;d=json.load(open('$O/${T}_128_tx${tx}.json'));print('128 tiles',$tx,'%.4e'%d['value'],'%.3f ms'%d['trace_ms_per_step'])"
done; done
for tx in 0 79; do
$B --n 256 --steps 2 --warmup 1 --tile-x $tx > $O/${T}_256_tx${tx}.json 2>> $O/${T}_err.log; python -c "import json;d=json.load(open('$O/${T}_256_tx${tx}.json'));print('256 tiles',$tx,'%.4e'%d['value'],'%.3f ms'%d['trace_ms_per_step'])"
done
tail -3 $O/${T}_err.log
