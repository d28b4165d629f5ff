set -x
python tests/ab_quad.py || exit 1
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["executed_steps_per_ray"])'
for lib in librmcrt_b200.so librmcrt_b200_q7.so librmcrt_b200_q8.so; do
for cfg in "--cells-per-lane 4" "--cells-per-lane 4 --coherence 2"; do
  echo "### $lib $cfg"; RMCRT_B200_LIB=$PWD/uintah_b200/lib/$lib $B $cfg 2>&1 | tail -1 | python -c "$P"
done; done
