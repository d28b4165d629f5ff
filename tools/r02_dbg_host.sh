cd $GRAFT_REPO_ROOT
python -m pytest tests/test_host_cpp.py -m gpu -q -k "radiometer_bounded_lhc or double_deterministic" 2>&1 | tail -3
RMCRT_PAGEABLE_HOST=1 python -m pytest tests/test_host_cpp.py -m gpu -q -k "radiometer_bounded_lhc" 2>&1 | tail -3
python -m pytest tests/test_host_cpp.py -m gpu -q 2>&1 | tail -3
