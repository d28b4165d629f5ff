# round 2, GPU visit 6: what the driver runs at round end (suite, smoke, bench, reference arm) + the ncu launch list of the bench command
T=r02v6
O=gpurun_out
python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $O/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"; cut -c1-300 $O/${T}_bench.json; tail -3 $O/${T}_bench.err
python bench.py --impl reference --steps 2 --warmup 3 > $O/${T}_ref.json 2>/dev/null; cut -c1-200 $O/${T}_ref.json
python bench.py --coherence 2 --no-sweeps --no-cpu-baseline > $O/${T}_bench_coh2.json 2>/dev/null; cut -c1-200 $O/${T}_bench_coh2.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sweeps > $O/${T}_b.log 2>&1
python tools/bench_configs.py --config C5 --steps 1 > $O/${T}_c5.json 2>/dev/null; cat $O/${T}_c5.json | cut -c1-300
