# round 2, GPU visit 59: counters of the default kernel with the hemisphere-major key (128^3), same metric list as r02v51_256_counters.csv
T=r02v59
O=gpurun_out
K='regex:divq_sorted_kernel<.*bool.0, .bool.0, .bool.0>'
M=l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__t_sector_hit_rate.pct,l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
timeout 100 ncu --metrics $M --clock-control none --kernel-name-base demangled -k "$K" -s 3 -c 1 --csv --log-file $O/${T}_128_counters.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sweeps > /dev/null 2>&1
cut -d, -f13- $O/${T}_128_counters.csv | tail -13
