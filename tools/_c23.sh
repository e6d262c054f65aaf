set -x
(time timeout 600 python -m pytest tests -m gpu -x -q) > gpurun_out/r1p_pytest.txt 2>&1; tail -4 gpurun_out/r1p_pytest.txt
timeout 300 python bench.py --n-patches 10000 --slides-per-step 64 --no-resnet --no-eager-baseline > gpurun_out/r1p_bench_n10k.json 2>/dev/null; cut -c1-120 gpurun_out/r1p_bench_n10k.json
timeout 300 python bench.py --n-patches 10000 --slides-per-step 64 --streams 1 --no-resnet --no-eager-baseline --no-cpu-baseline > gpurun_out/r1p_bench_n10k_s1.json 2>/dev/null; cut -c1-120 gpurun_out/r1p_bench_n10k_s1.json
timeout 900 python bench.py > gpurun_out/r1p_bench.json 2> gpurun_out/r1p_bench.err; tail -c 300 gpurun_out/r1p_bench.err; cut -c1-120 gpurun_out/r1p_bench.json
