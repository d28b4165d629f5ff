# round 2, GPU visit 24 (2 GPUs): multi-device tests and the N=2 bench on the sorted defaults
T=r02v24
O=gpurun_out
python -m pytest tests/test_gpu_multi.py tests/test_host_cpp.py -m gpu -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/${T}_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/${T}_bench_n2.json 2> $O/${T}_bench_n2.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02v24_bench_n2.json'))
print('value',d['value'],'ms',d['ms_per_step'],'trace',d['trace_ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'])
print({k:v['ray_cell_steps_per_s'] for k,v in d['by_coherence'].items()})
print({k:(v['ray_cell_steps_per_s'],v['ms_per_solve']) for k,v in d['configs'].items()})
PY
tail -3 $O/${T}_bench_n2.err
for nd in 1 2; do uintah_b200/host/rmcrt_test tests/decks/bm1_1L_float_128_perf.ups --timesteps 3 --devices $nd 2>&1 | grep -E "Used|REPORT" | tail -2; done
