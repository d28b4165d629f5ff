T=${1:-s9i}
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/${T}_bench.json
python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/${T}_ref.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:divq_quad_kernel -s 3 -c 1 -f -o gpurun_out/${T}_divq_quad_128 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu.log 2>&1
python tools/bench_configs.py > gpurun_out/${T}_configs.json 2>/dev/null; cut -c1-100,190-330 gpurun_out/${T}_configs.json
