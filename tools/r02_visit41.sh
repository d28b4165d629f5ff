# round 2, GPU visit 41: tile groups only up to 16 rays per cell, G == 1 fast paths -- headline back?
T=r02v41
O=gpurun_out
B="python bench.py --no-cpu-baseline --no-sweeps"
for rays in 100 25 16 10 5; do
$B --steps 8 --rays $rays > $O/${T}_r${rays}.json 2>> $O/${T}_err.log; python -c "import json;d=json.load(open('$O/${T}_r${rays}.json'));print('rays',$rays,'%.4e'%d['value'],'%.3f ms'%d['trace_ms_per_step'])"
done
$B --n 256 --steps 2 --warmup 1 > $O/${T}_256.json 2>> $O/${T}_err.log; python -c "import json;d=json.load(open('$O/${T}_256.json'));print('256','%.4e'%d['value'],d['trace_ms_per_step'])"
python tools/bench_configs.py --config C4 C5 --steps 2 2>/dev/null | cut -c1-12,150-300
python -m pytest tests/test_gpu_indep.py -m gpu -q -x 2>&1 | tail -2
tail -3 $O/${T}_err.log
