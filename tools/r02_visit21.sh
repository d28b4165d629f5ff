# round 2, GPU visit 21: whole GPU suite on the sorted / blocked defaults (single level + Data Onion), smoke
T=r02v21
O=gpurun_out
python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $O/${T}_pytest.log
python -m pytest tests/test_gpu_atsize.py -m gpu -q -s > $O/${T}_atsize.log 2>&1; grep -E "f3|L2|passed|failed|centre" $O/${T}_atsize.log | tail -30
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
