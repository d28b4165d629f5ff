# round 2, GPU visit 30: compute-sanitizer racecheck (shared-memory hazards of the sort phases) and memcheck of a few more paths
T=r02v30
O=gpurun_out
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_indep.py -m gpu -q -x -k "sorted_kernel_is_independent or sorted_data_onion_three_levels or (blocked_march_matches and plates and 3)" > $O/${T}_racecheck.log 2>&1; echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|passed|failed|hazard" $O/${T}_racecheck.log | head -12
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_atsize.py -m gpu -q -x -k "data_onion or flux" > $O/${T}_memcheck2.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid" $O/${T}_memcheck2.log | head -8
cp $O/r02v29_memcheck.log $O/${T}_memcheck1.log 2>/dev/null
