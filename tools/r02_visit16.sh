# round 2, GPU visit 16: ray-length estimate with mirror faces and the intensity threshold -- C5 and the reflective parity cases
T=r02v16
O=gpurun_out
python -m pytest tests/test_gpu_indep.py tests/test_gpu_parity.py tests/test_bench_contract.py -m gpu -q -x > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/${T}_pytest.log
python tools/bench_configs.py --config C5 --steps 2 > $O/${T}_c5.json 2> $O/${T}_c5.err; cut -c1-600 $O/${T}_c5.json
python bench.py --no-cpu-baseline --no-sweeps --steps 5 > $O/${T}_128.json 2>> $O/${T}_err.log; python -c "import json;d=json.load(open('$O/${T}_128.json'));print('128',d['value'],d['trace_ms_per_step'])"
tail -3 $O/${T}_c5.err
