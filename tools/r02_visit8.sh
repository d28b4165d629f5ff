# round 2, 8 GPUs: the driver's scaling run at N=8 (and N=4), multi-device tests
T=r02v8
O=gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -2
for N in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > $O/${T}_bench_n$N.json 2> $O/${T}_bench_n$N.err; echo "bench N=$N rc=$?"
python - <<PY
import json
txt=open('$O/${T}_bench_n$N.json').read()
d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
print('N=$N value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'])
print({k:v['ray_cell_steps_per_s'] for k,v in d['by_coherence'].items()})
print({k:(v['ray_cell_steps_per_s'],v['ms_per_solve']) for k,v in d['configs'].items()})
PY
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --impl reference --gpus 8 --steps 2 --warmup 3 2>/dev/null | cut -c1-300
uintah_b200/host/rmcrt_test tests/decks/bm1_1L_float_128_perf.ups --timesteps 3 --devices 8 2>&1 | grep -E "Used" | tail -1
