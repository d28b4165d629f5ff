# round 2, GPU visit 54: full ncu captures of the final Data-Onion and RESORT kernels (compact tiles, rounds)
T=r02v54
O=gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:divq_sorted_kernel<.*bool.1, .bool.0, .bool.0>" -s 8 -c 1 -f -o $O/${T}_c4_divq_sorted_onion python tools/bench_configs.py --config C4 --steps 1 > $O/${T}_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:divq_sorted_kernel<.*bool.0, .bool.1, .bool.0>" -s 1 -c 1 -f -o $O/${T}_c5_divq_sorted_resort_128 python tools/bench_configs.py --config C5 --n 128 --steps 1 > $O/${T}_ncu4.log 2>&1
ls -la $O/${T}_*.ncu-rep
