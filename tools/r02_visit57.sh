# round 2, GPU visit 57: sort key A/B -- (length | octant) [shipped] against octant-major and hemisphere-major keys
T=r02v57
O=gpurun_out
B="python bench.py --no-cpu-baseline --no-sweeps --steps 8"
for v in 0 1 2; do
  L=$PWD/uintah_b200/lib/librmcrt_b200.so; [ $v != 0 ] && L=$PWD/uintah_b200/lib/librmcrt_b200_kv$v.so
  RMCRT_B200_LIB=$L $B > $O/${T}_kv$v.json 2>> $O/${T}_err.log
  python -c "import json;d=json.load(open('$O/${T}_kv$v.json'));print('kv$v','%.4e'%d['value'],'%.3f ms'%d['trace_ms_per_step'])"
done
tail -2 $O/${T}_err.log
