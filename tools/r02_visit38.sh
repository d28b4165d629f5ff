# round 2, GPU visit 38: boundary-flux rays on the blocked copy -- flux tests, C5
T=r02v38
O=gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_atsize.py tests/test_gpu_features.py tests/test_host_cpp.py -m gpu -q -x > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${T}_pytest.log
python tools/bench_configs.py --config C5 --steps 2 > $O/${T}_c5.json 2> $O/${T}_c5.err; cut -c1-420 $O/${T}_c5.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/${T}_c5_launches.csv python tools/bench_configs.py --config C5 --steps 1 > /dev/null 2>&1
grep -E "flux_trace|divq_sorted" $O/${T}_c5_launches.csv | awk -F'","' '{print $5, $NF}' | cut -c1-60,200-260 | tail -6
tail -3 $O/${T}_c5.err
