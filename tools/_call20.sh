set -x
(time timeout 600 python -m pytest tests -m gpu -x -q) > gpurun_out/c20_pytest.txt 2>&1; tail -5 gpurun_out/c20_pytest.txt
timeout 300 python tools/gpu_diag.py --step train_fused > gpurun_out/c20_train_fused.txt 2>&1; tail -1 gpurun_out/c20_train_fused.txt
timeout 300 python tools/abtime.py toad_b200/libtoad_b200.so > gpurun_out/c20_ab.json 2>&1; cat gpurun_out/c20_ab.json
timeout 300 python tools/resnet_sweep.py 256 > gpurun_out/c20_resnet.json 2>&1; cat gpurun_out/c20_resnet.json
