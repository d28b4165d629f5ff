# round 2, GPU visit 40: tile groups (G tiles sorted together when a cell has few rays) -- parity tests, few-ray sweep, headline unchanged
T=r02v40
O=gpurun_out
python -m pytest tests/test_gpu_indep.py tests/test_gpu_parity.py tests/test_gpu_atsize.py -m gpu -q -x > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/${T}_pytest.log
B="python bench.py --no-cpu-baseline --no-sweeps --steps 5"
for rays in 3 5 10 25 50 100; do for bm in 3 2; do
$B --rays $rays --blocked-march $bm > $O/${T}_r${rays}_bm${bm}.json 2>> $O/${T}_err.log; python -c "import json;d=json.load(open('$O/${T}_r${rays}_bm${bm}.json'));print('rays',$rays,'bm',$bm,'%.3e'%d['value'],'%.3f ms'%d['trace_ms_per_step'])"
done; done
$B --n 256 --steps 2 --warmup 1 > $O/${T}_256.json 2>> $O/${T}_err.log; python -c "import json;d=json.load(open('$O/${T}_256.json'));print('256','%.4e'%d['value'],d['trace_ms_per_step'])"
tail -3 $O/${T}_err.log
