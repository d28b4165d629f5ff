# one GPU visit: parity tests, bench, reference arm, launch list, full ncu capture of the hot kernel (tag = $1)
T=${1:-s9}
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/${T}_bench.json
python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/${T}_ref.json 2>/dev/null; cut -c1-300 gpurun_out/${T}_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:divq_quad_kernel -s 3 -c 1 -f -o gpurun_out/${T}_divq_quad_128 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu.log 2>&1
python tools/bench_configs.py > gpurun_out/${T}_configs.json 2> gpurun_out/${T}_configs.err; cat gpurun_out/${T}_configs.json | cut -c1-300
