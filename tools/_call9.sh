set -x
(time timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_gpu_kernels.py -m gpu -x -q) > gpurun_out/c9_pytest.txt 2>&1; tail -5 gpurun_out/c9_pytest.txt
for f in 0 32 64 96; do timeout 300 python tools/abtime.py toad_b200/libtoad_b200.so $f > gpurun_out/c9_ab_f$f.json 2>&1; cat gpurun_out/c9_ab_f$f.json; done
timeout 600 python bench.py --no-resnet > gpurun_out/c9_bench.json 2> gpurun_out/c9_bench.err; tail -c 300 gpurun_out/c9_bench.err
