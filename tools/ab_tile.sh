B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["executed_steps_per_ray"])'
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "four_cells or kernel_choice" 2>&1 | tail -2
for cfg in "--tile-x 32" "--tile-x 16" "--tile-x 8" "--tile-x 8 --coherence 1" "--tile-x 8 --n 256" "--tile-x 16 --n 256" "--tile-x 8 --warps 2" "--tile-x 8 --half-records 1"; do
  echo "### $cfg"; $B $cfg 2>&1 | tail -1 | python -c "$P"
done
