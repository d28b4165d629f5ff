# round 2, GPU visit 15: divq_sorted_kernel as the library default -- whole GPU suite, smoke, driver-format bench, full ncu capture + launch list
T=r02v15
O=gpurun_out
python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $O/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"; cut -c1-400 $O/${T}_bench.json
ncu --set full --clock-control none --import-source on -k regex:divq_sorted_kernel -s 3 -c 1 -f -o $O/${T}_divq_sorted_128 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sweeps > $O/${T}_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sweeps > $O/${T}_b.log 2>&1
tail -3 $O/${T}_bench.err
