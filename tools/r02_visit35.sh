# round 2, GPU visit 35: persisting L2 window on / off at 256^3 (72-MB blocked copy) and 128^3
T=r02v35
O=gpurun_out
B="python bench.py --no-cpu-baseline --no-sweeps"
for p in 1 0; do
$B --n 256 --steps 3 --warmup 1 --l2persist $p > $O/${T}_256_p${p}.json 2>> $O/${T}_err.log; python -c "import json;d=json.load(open('$O/${T}_256_p${p}.json'));print('256 l2persist',$p,'%.4e'%d['value'],d['trace_ms_per_step'])"
$B --steps 5 --l2persist $p > $O/${T}_128_p${p}.json 2>> $O/${T}_err.log; python -c "import json;d=json.load(open('$O/${T}_128_p${p}.json'));print('128 l2persist',$p,'%.4e'%d['value'],d['trace_ms_per_step'])"
done
python - <<'PY'
import torch
p=torch.cuda.get_device_properties(0)
print("L2", p.L2_cache_size)
import ctypes
rt=ctypes.CDLL("libcudart.so")
v=ctypes.c_int()
for name,attr in (("maxPersistingL2",108),("maxAccessPolicyWindow",109)):
    rt.cudaDeviceGetAttribute(ctypes.byref(v), attr, 0); print(name, v.value)
PY
tail -3 $O/${T}_err.log
