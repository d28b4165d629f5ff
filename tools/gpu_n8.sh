N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/s9_bench_n$N.json 2> gpurun_out/s9_bench_n$N.err; tail -1 gpurun_out/s9_bench_n$N.json | cut -c1-330
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 tools/bench_configs.py --config C4 > gpurun_out/s9_c4_n$N.json 2> gpurun_out/s9_c4_n$N.err; tail -1 gpurun_out/s9_c4_n$N.json | cut -c1-400
