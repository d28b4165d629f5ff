# round 2, GPU visit 2: at-size parity, whole GPU suite with the conformant default, bench with by_coherence / configs
T=r02v2
O=gpurun_out
python -m pytest tests/test_gpu_atsize.py -q -s -k "not boundary_flux" > $O/${T}_atsize.log 2>&1; echo "atsize rc=$?"; grep -E "passed|failed|error" $O/${T}_atsize.log | tail -3
python -m pytest tests -m gpu -q --deselect tests/test_gpu_atsize.py > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $O/${T}_pytest.log
python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"; cut -c1-600 $O/${T}_bench.json; tail -5 $O/${T}_bench.err
python bench.py --impl reference --steps 2 --warmup 3 > $O/${T}_ref.json 2>/dev/null; cut -c1-300 $O/${T}_ref.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
