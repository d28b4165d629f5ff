# round 2, GPU visit 31 (4 GPUs): the driver's N=4 bench command on the shipped kernels
T=r02v34
O=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 10 --warmup 3 > $O/${T}_bench_n4.json 2> $O/${T}_bench_n4.err; echo "bench rc=$?"
python - <<'PY'
import json
txt=open('gpurun_out/r02v34_bench_n4.json').read()
d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'trace',d['trace_ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'])
print({k:v['ray_cell_steps_per_s'] for k,v in d['by_coherence'].items()})
print({k:(v['ray_cell_steps_per_s'],v['ms_per_solve']) for k,v in d['configs'].items()})
PY
tail -3 $O/${T}_bench_n4.err
