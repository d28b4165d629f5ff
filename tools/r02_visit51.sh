# round 2, GPU visit 51: final code -- whole GPU suite, smoke, driver-format bench, reference arm, launch list, ncu captures
T=r02v51
O=gpurun_out
python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"; cut -c1-250 $O/${T}_bench.json
python bench.py --impl reference --steps 2 --warmup 3 > $O/${T}_ref.json 2>/dev/null; cut -c1-160 $O/${T}_ref.json
K='regex:divq_sorted_kernel<.*bool.0, .bool.0, .bool.0>'
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "$K" -s 3 -c 1 -f -o $O/${T}_divq_sorted_128 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sweeps > $O/${T}_ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sweeps > $O/${T}_b.log 2>&1
M=l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__t_sector_hit_rate.pct,l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
ncu --metrics $M --clock-control none --kernel-name-base demangled -k "$K" -s 1 -c 1 --csv --log-file $O/${T}_256_counters.csv python bench.py --n 256 --steps 1 --warmup 1 --no-cpu-baseline --no-sweeps > /dev/null 2>&1
tail -2 $O/${T}_ncu1.log | cut -c1-150
tail -3 $O/${T}_bench.err
