set -x
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/c12_pytest.txt 2>&1; tail -8 gpurun_out/c12_pytest.txt
timeout 300 python tools/topk_time.py > gpurun_out/c12_topk.json 2>&1; cat gpurun_out/c12_topk.json
timeout 300 python tools/gpu_diag.py --step train_fused > gpurun_out/c12_train_fused.txt 2>&1; tail -5 gpurun_out/c12_train_fused.txt
timeout 300 python tools/gpu_diag.py --step train_breakdown > gpurun_out/c12_train_breakdown.txt 2>&1; tail -5 gpurun_out/c12_train_breakdown.txt
