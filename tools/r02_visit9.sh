# round 2, GPU visit 9: A/B of north_star's shared-memory staging of the origin block (generic march, independent rays, 128^3 x 100 rays)
T=r02v9
O=gpurun_out
python -m pytest tests/test_gpu_indep.py -m gpu -q -k brick 2>&1 | tail -3
B="python bench.py --no-cpu-baseline --no-sweeps --steps 5 --generic-march 1"
for h in 0 2 4 6 8; do
$B --origin-brick-halo $h > $O/${T}_brick_h$h.json 2>> $O/${T}_err.log; python -c "import json;d=json.load(open('$O/${T}_brick_h$h.json'));print('halo',$h,d['value'],d['trace_ms_per_step'])"
done
ncu --metrics l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:divq_brick_kernel -s 3 -c 1 --csv --log-file $O/${T}_brick_h6_counters.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sweeps --generic-march 1 --origin-brick-halo 6 > /dev/null 2>&1
ncu --metrics l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:divq_kernel -s 3 -c 1 --csv --log-file $O/${T}_generic_counters.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sweeps --generic-march 1 > /dev/null 2>&1
tail -8 $O/${T}_brick_h6_counters.csv | cut -c1-200; tail -8 $O/${T}_generic_counters.csv | cut -c1-200
tail -3 $O/${T}_err.log
