python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02v52_bench_n2.json 2> gpurun_out/r02v52_bench_n2.err
python - <<'PY'
import json
txt=open('gpurun_out/r02v52_bench_n2.json').read()
d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
print({k:(v['ray_cell_steps_per_s'],v['ms_per_solve']) for k,v in d['configs'].items()})
PY
