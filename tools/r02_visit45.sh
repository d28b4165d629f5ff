# round 2, GPU visit 45 (2 GPUs): multi-device tests and the N=2 bench on the final code
T=r02v45
O=gpurun_out
python -m pytest tests/test_gpu_multi.py tests/test_host_cpp.py -m gpu -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${T}_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/${T}_bench_n2.json 2> $O/${T}_bench_n2.err; echo "bench rc=$?"
python - <<'PY'
import json
txt=open('gpurun_out/r02v45_bench_n2.json').read()
d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
print({k:(v['ray_cell_steps_per_s'],v['ms_per_solve']) for k,v in d['configs'].items()})
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>/dev/null | grep '^{' | cut -c1-200
