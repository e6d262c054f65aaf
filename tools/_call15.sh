set -x
(time timeout 600 python -m pytest tests/test_gpu_backward.py tests/test_gpu_train.py tests/test_gpu_dropout.py -m gpu -x -q) > gpurun_out/c15_pytest.txt 2>&1; tail -6 gpurun_out/c15_pytest.txt
timeout 300 python tools/gpu_diag.py --step train_fused > gpurun_out/c15_train_fused.txt 2>&1; tail -2 gpurun_out/c15_train_fused.txt
