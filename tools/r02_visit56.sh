# round 2, GPU visit 56: the library as rebuilt after the refillLanes clamp -- whole GPU suite, smoke, driver-format bench
T=r02v56
O=gpurun_out
python -m pytest tests -m gpu -q -x > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"; cut -c1-250 $O/${T}_bench.json
tail -2 $O/${T}_bench.err
