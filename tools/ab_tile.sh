B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["e2e"]["value"])'
for cfg in "--tile-x 16" "--tile-x 8" "--tile-x 32" "--tile-x 16 --warps 2" "--tile-x 8 --n 256" "--tile-x 16 --n 256"; do
  echo "### $cfg"; $B $cfg 2>&1 | tail -1 | python -c "$P"
done
