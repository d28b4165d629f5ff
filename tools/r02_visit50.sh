# round 2, GPU visit 50: compact origin tiles (4 x 4 x 2) as the rule -- parity suites, every config
T=r02v50
O=gpurun_out
python -m pytest tests/test_gpu_indep.py tests/test_gpu_parity.py tests/test_gpu_atsize.py tests/test_gpu_features.py tests/test_host_cpp.py -m gpu -q -x > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${T}_pytest.log
B="python bench.py --no-cpu-baseline --no-sweeps"
$B --steps 10 > $O/${T}_128.json 2>> $O/${T}_err.log; python -c "import json;d=json.load(open('$O/${T}_128.json'));print('128','%.4e'%d['value'],'%.3f ms'%d['trace_ms_per_step'],'e2e %.4e'%d['e2e']['value'])"
$B --n 256 --steps 2 --warmup 1 > $O/${T}_256.json 2>> $O/${T}_err.log; python -c "import json;d=json.load(open('$O/${T}_256.json'));print('256','%.4e'%d['value'],'%.3f ms'%d['trace_ms_per_step'])"
for rays in 5 25; do $B --steps 5 --rays $rays > $O/${T}_r$rays.json 2>> $O/${T}_err.log; python -c "import json;d=json.load(open('$O/${T}_r$rays.json'));print('rays',$rays,'%.4e'%d['value'])"; done
python tools/bench_configs.py --config C4 C5 --steps 2 2>/dev/null | cut -c1-12,150-300
tail -3 $O/${T}_err.log
