T=${1:-s9i}
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/${T}_bench.json
ncu --set full --clock-control none --import-source on -k regex:divq_quad_kernel -s 2 -c 1 -f -o gpurun_out/${T}_divq_quad_onion_c4 python tools/bench_configs.py --config C4 --steps 1 > gpurun_out/${T}_ncu_c4.log 2>&1; tail -2 gpurun_out/${T}_ncu_c4.log | cut -c1-200
