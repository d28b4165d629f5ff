# round 2, GPU visit 29: compute-sanitizer memcheck over the tests of the sorted / blocked kernels (single level, RESORT, Data Onion 2 and 3 levels)
T=r02v29
O=gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_indep.py -m gpu -q -x -k "sorted or blocked_march_matches" > $O/${T}_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" $O/${T}_memcheck.log | head -20
