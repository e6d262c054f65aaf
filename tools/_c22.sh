set -x
timeout 300 python bench.py --n-patches 10000 --slides-per-step 64 --no-resnet --no-eager-baseline > gpurun_out/r1o_bench_n10k.json 2>/dev/null; cut -c1-160 gpurun_out/r1o_bench_n10k.json
timeout 300 python bench.py --n-patches 10000 --slides-per-step 64 --streams 6 --no-resnet --no-eager-baseline --no-cpu-baseline > gpurun_out/r1o_bench_n10k_s6.json 2>/dev/null; cut -c1-160 gpurun_out/r1o_bench_n10k_s6.json
