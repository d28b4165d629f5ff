# round 2, GPU visit 58: hemisphere-major sort key as the default -- whole GPU suite, smoke, driver-format bench (with the 256^3 / C4 / C5 configs)
T=r02v58
O=gpurun_out
python -m pytest tests -m gpu -q -x > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"; cut -c1-250 $O/${T}_bench.json
python -c "
import json;d=json.load(open('$O/${T}_bench.json'))
for k,v in d['configs'].items(): print(k,'%.4e'%v['ray_cell_steps_per_s'])"
