# round 2, GPU visit 26: full ncu captures of the single-level default kernel (the plain instantiation, not the gated-off RESORT launch)
T=r02v26
O=gpurun_out
K='regex:divq_sorted_kernel<.*bool.0, .bool.0>'
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "$K" -s 3 -c 1 -f -o $O/${T}_divq_sorted_128 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sweeps > $O/${T}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "$K" -s 1 -c 1 -f -o $O/${T}_divq_sorted_256_rec4 python bench.py --n 256 --steps 1 --warmup 1 --no-cpu-baseline --no-sweeps > $O/${T}_ncu2.log 2>&1
tail -n 2 $O/${T}_ncu1.log | cut -c1-200
