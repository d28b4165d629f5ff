set -x
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
for cfg in "" "--coherence 2" "--half-records 1" "--coherence 2 --half-records 1" "--n 256" "--n 256 --coherence 2"; do
  echo "### $cfg"
  $B $cfg 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['executed_steps_per_ray'])"
done
