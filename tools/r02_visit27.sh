# round 2, GPU visit 27: N-level Data Onion on blocked boxes (march_onion_blk) -- parity tests, C4
T=r02v27
O=gpurun_out
python -m pytest tests/test_gpu_indep.py tests/test_gpu_parity.py tests/test_gpu_atsize.py tests/test_gpu_features.py -m gpu -q -x > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $O/${T}_pytest.log
python tools/bench_configs.py --config C4 --steps 2 > $O/${T}_c4.json 2> $O/${T}_c4.err; cut -c1-420 $O/${T}_c4.json
tail -3 $O/${T}_c4.err
