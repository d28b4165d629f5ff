set -x
RMCRT_B200_LIB=$PWD/uintah_b200/lib/librmcrt_b200_prev.so python tests/ab_dump.py gpurun_out/ab_prev.npz || exit 1
python tests/ab_dump.py gpurun_out/ab_new.npz || exit 1
python tests/ab_dump.py --compare gpurun_out/ab_prev.npz gpurun_out/ab_new.npz
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["executed_steps_per_ray"])'
for cfg in "" "--n 256"; do
  echo "### prev $cfg"; RMCRT_B200_LIB=$PWD/uintah_b200/lib/librmcrt_b200_prev.so $B $cfg 2>&1 | tail -1 | python -c "$P"
  echo "### new $cfg"; $B $cfg 2>&1 | tail -1 | python -c "$P"
done
