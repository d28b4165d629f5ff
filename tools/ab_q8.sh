python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "four_cells or kernel_choice or monte_carlo or reproducible or two_cells or isothermal" 2>&1 | tail -2
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["e2e"]["value"])'
for lib in librmcrt_b200.so librmcrt_b200_q8.so; do for cfg in "" "--n 256"; do
  echo "### $lib $cfg"; RMCRT_B200_LIB=$PWD/uintah_b200/lib/$lib $B $cfg 2>&1 | tail -1 | python -c "$P"
done; done
