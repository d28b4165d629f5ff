# round 2, GPU visit 25: driver-format bench + reference arm + launch list + full ncu captures of the three default kernels
T=r02v25
O=gpurun_out
python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"; cut -c1-300 $O/${T}_bench.json
python bench.py --impl reference --steps 2 --warmup 3 > $O/${T}_ref.json 2>/dev/null; cut -c1-200 $O/${T}_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sweeps > $O/${T}_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:divq_sorted_kernel -s 3 -c 1 -f -o $O/${T}_divq_sorted_128 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sweeps > $O/${T}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:divq_sorted_kernel -s 3 -c 1 -f -o $O/${T}_divq_sorted_256_rec4 python bench.py --n 256 --steps 1 --warmup 1 --no-cpu-baseline --no-sweeps > $O/${T}_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:divq_sorted_kernel -s 8 -c 1 -f -o $O/${T}_c4_divq_sorted_onion python tools/bench_configs.py --config C4 --steps 1 > $O/${T}_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:divq_sorted_kernel -s 3 -c 1 -f -o $O/${T}_c5_divq_sorted_resort_128 python tools/bench_configs.py --config C5 --n 128 --steps 1 > $O/${T}_ncu4.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/${T}_c45_launches.csv python tools/bench_configs.py --config C4 C5 --steps 1 > $O/${T}_b2.log 2>&1
tail -3 $O/${T}_bench.err
