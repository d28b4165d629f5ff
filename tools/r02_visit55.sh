# round 2, GPU visit 55: resident blocks per SM again, now with compact tiles (5 blocks need <= 51 registers)
T=r02v55
O=gpurun_out
B="python bench.py --no-cpu-baseline --no-sweeps --steps 8"
$B > $O/${T}_k4.json 2>> $O/${T}_err.log; python -c "import json;d=json.load(open('$O/${T}_k4.json'));print('k4','%.4e'%d['value'],'%.3f ms'%d['trace_ms_per_step'])"
$B --max-blocks-per-sm 3 > $O/${T}_k3.json 2>> $O/${T}_err.log; python -c "import json;d=json.load(open('$O/${T}_k3.json'));print('k3','%.4e'%d['value'],'%.3f ms'%d['trace_ms_per_step'])"
RMCRT_B200_LIB=$PWD/uintah_b200/lib/librmcrt_b200_k5.so $B > $O/${T}_k5.json 2>> $O/${T}_err.log; python -c "import json;d=json.load(open('$O/${T}_k5.json'));print('k5','%.4e'%d['value'],'%.3f ms'%d['trace_ms_per_step'])"
RMCRT_B200_LIB=$PWD/uintah_b200/lib/librmcrt_b200_k5.so $B --n 256 --steps 2 --warmup 1 > $O/${T}_k5_256.json 2>> $O/${T}_err.log; python -c "import json;d=json.load(open('$O/${T}_k5_256.json'));print('k5 256','%.4e'%d['value'],'%.3f ms'%d['trace_ms_per_step'])"
tail -3 $O/${T}_err.log
