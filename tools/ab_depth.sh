set -x
RMCRT_B200_LIB=$PWD/uintah_b200/lib/librmcrt_b200_prev.so python tests/ab_dump.py gpurun_out/ab_prev.npz || exit 1
python tests/ab_dump.py gpurun_out/ab_new.npz || exit 1
python tests/ab_dump.py --compare gpurun_out/ab_prev.npz gpurun_out/ab_new.npz
RMCRT_B200_LIB=$PWD/uintah_b200/lib/librmcrt_b200_prev.so python tools/bench_configs.py --config C5 | cut -c1-60,190-330
python tools/bench_configs.py --config C5 | cut -c1-60,190-330
