# round 2, GPU visit 4: whole GPU suite (new: coarsening logic, coarsenExtraCells, fixed ROI, C5 at size), C++ plugin timing
T=r02v4
O=gpurun_out
python -m pytest tests -m gpu -q -x > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $O/${T}_pytest.log
grep -E '^\.?F?\{"what": "c5' $O/${T}_pytest.log | cut -c1-300
uintah_b200/host/rmcrt_test tests/decks/bm1_1L_float_128_perf.ups --timesteps 3 2>&1 | grep -i -E "report|solve|ms|steps" | tail -8
