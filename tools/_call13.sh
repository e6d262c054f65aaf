set -x
(time timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_eval.py tests/test_gpu_forward.py -m gpu -x -q) > gpurun_out/c13_pytest.txt 2>&1; tail -8 gpurun_out/c13_pytest.txt
timeout 300 python tools/topk_time.py > gpurun_out/c13_topk.json 2>&1; cat gpurun_out/c13_topk.json
timeout 300 python tools/topk_time.py 50000 > gpurun_out/c13_topk50k.json 2>&1; cat gpurun_out/c13_topk50k.json
