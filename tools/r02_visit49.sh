# round 2, GPU visit 49: compact tiles aligned to the 2 x 2 x 2 blocks (guard of 3 cells) against the unaligned copy, same box
T=r02v49
O=gpurun_out
B="python bench.py --no-cpu-baseline --no-sweeps"
for lib in librmcrt_b200_prev.so librmcrt_b200.so; do for tx in 80 81; do
RMCRT_B200_LIB=$PWD/uintah_b200/lib/$lib $B --steps 8 --tile-x $tx > $O/${T}_128.json 2>> $O/${T}_err.log; python -c "import json;d=json.load(open('$O/${T}_128.json'));print('$lib 128 shape',$tx,'%.4e'%d['value'],'%.3f ms'%d['trace_ms_per_step'])"
done
RMCRT_B200_LIB=$PWD/uintah_b200/lib/$lib $B --n 256 --steps 2 --warmup 1 --tile-x 80 > $O/${T}_256.json 2>> $O/${T}_err.log; python -c "import json;d=json.load(open('$O/${T}_256.json'));print('$lib 256 shape 80','%.4e'%d['value'],'%.3f ms'%d['trace_ms_per_step'])"
done
python -m pytest tests/test_gpu_indep.py -m gpu -q -x 2>&1 | tail -2
tail -3 $O/${T}_err.log
