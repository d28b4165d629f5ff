# round 2, GPU visit 37: streaming loads / stores for the once-per-cell traffic of the sorted kernel
T=r02v37
O=gpurun_out
B="python bench.py --no-cpu-baseline --no-sweeps"
$B --steps 10 > $O/${T}_128.json 2>> $O/${T}_err.log; python -c "import json;d=json.load(open('$O/${T}_128.json'));print('128','%.4e'%d['value'],d['trace_ms_per_step'],'e2e %.4e'%d['e2e']['value'])"
$B --n 256 --steps 3 --warmup 1 > $O/${T}_256.json 2>> $O/${T}_err.log; python -c "import json;d=json.load(open('$O/${T}_256.json'));print('256','%.4e'%d['value'],d['trace_ms_per_step'])"
python tools/bench_configs.py --config C4 --steps 2 2>/dev/null | cut -c150-330
python -m pytest tests/test_gpu_indep.py tests/test_gpu_atsize.py -m gpu -q -x 2>&1 | tail -2
tail -3 $O/${T}_err.log
