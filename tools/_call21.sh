set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r1n_bench_8gpu.json 2> gpurun_out/r1n_bench_8gpu.err; tail -c 300 gpurun_out/r1n_bench_8gpu.err; cut -c1-250 gpurun_out/r1n_bench_8gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r1n_bench_4gpu.json 2>/dev/null; cut -c1-250 gpurun_out/r1n_bench_4gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 tools/train_synthetic.py --slides 512 > gpurun_out/r1n_train_8gpu.txt 2>&1; tail -2 gpurun_out/r1n_train_8gpu.txt
