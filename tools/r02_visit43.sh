# round 2, GPU visit 43: whole GPU suite + smoke + driver-format bench + reference arm on the final code
T=r02v43
O=gpurun_out
python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"; cut -c1-250 $O/${T}_bench.json
python bench.py --impl reference --steps 2 --warmup 3 > $O/${T}_ref.json 2>/dev/null; cut -c1-160 $O/${T}_ref.json
python bench.py --no-cpu-baseline --no-sweeps --steps 5 --rays 5 > $O/${T}_rays5.json 2>> $O/${T}_err.log; python -c "import json;d=json.load(open('$O/${T}_rays5.json'));print('rays 5 default','%.3e'%d['value'],d['trace_ms_per_step'])"
tail -3 $O/${T}_bench.err
