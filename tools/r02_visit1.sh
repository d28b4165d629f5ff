# round 2, GPU visit 1: where does the conformant sampling (rayCoherence 0, Philox keyed on (cell, ray, timestep)) stand, and the ncu
# captures round 1 left open (C5 kernels)
T=r02v1
O=gpurun_out
B="python bench.py --no-cpu-baseline --steps 5"
$B --coherence 0                          > $O/${T}_coh0_128_rec8.json 2> $O/${T}_err.log; cut -c1-200 $O/${T}_coh0_128_rec8.json
$B --coherence 0 --half-records 1         > $O/${T}_coh0_128_rec4.json 2>> $O/${T}_err.log; cut -c1-120 $O/${T}_coh0_128_rec4.json
$B --coherence 0 --warps 8                > $O/${T}_coh0_128_w8.json 2>> $O/${T}_err.log; cut -c1-120 $O/${T}_coh0_128_w8.json
$B --coherence 0 --warps 2                > $O/${T}_coh0_128_w2.json 2>> $O/${T}_err.log; cut -c1-120 $O/${T}_coh0_128_w2.json
$B --coherence 0 --l2persist 0            > $O/${T}_coh0_128_nol2p.json 2>> $O/${T}_err.log; cut -c1-120 $O/${T}_coh0_128_nol2p.json
$B --coherence 0 --n 256 --steps 2        > $O/${T}_coh0_256_rec4.json 2>> $O/${T}_err.log; cut -c1-120 $O/${T}_coh0_256_rec4.json
$B --coherence 0 --n 256 --steps 2 --half-records 0 > $O/${T}_coh0_256_rec8.json 2>> $O/${T}_err.log; cut -c1-120 $O/${T}_coh0_256_rec8.json
$B --coherence 1 --cells-per-lane 1       > $O/${T}_coh1_cpl1_128.json 2>> $O/${T}_err.log; cut -c1-120 $O/${T}_coh1_cpl1_128.json
$B --coherence 1                          > $O/${T}_coh1_cpl4_128.json 2>> $O/${T}_err.log; cut -c1-120 $O/${T}_coh1_cpl4_128.json
$B --coherence 2                          > $O/${T}_coh2_cpl4_128.json 2>> $O/${T}_err.log; cut -c1-120 $O/${T}_coh2_cpl4_128.json
ncu --set full --clock-control none --import-source on -k regex:divq_kernel -s 3 -c 1 -f -o $O/${T}_divq_coh0_128 python bench.py --coherence 0 --steps 1 --warmup 3 --no-cpu-baseline > $O/${T}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:divq_kernel -s 3 -c 1 -f -o $O/${T}_divq_coh0_128_rec4 python bench.py --coherence 0 --half-records 1 --steps 1 --warmup 3 --no-cpu-baseline > $O/${T}_ncu2.log 2>&1
# C5 kernels (reflective-wall pair kernel and the boundary-flux trace)
ncu --set full --clock-control none --import-source on -k regex:divq_dual_kernel -s 1 -c 1 -f -o $O/${T}_c5_divq_dual_rejoin python tools/bench_configs.py --config C5 --steps 1 > $O/${T}_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:flux_trace_kernel -s 1 -c 1 -f -o $O/${T}_c5_flux_trace python tools/bench_configs.py --config C5 --steps 1 > $O/${T}_ncu4.log 2>&1
ls -la $O | tail -20
