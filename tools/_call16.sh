set -x
(time timeout 600 python -m pytest tests -m gpu -x -q) > gpurun_out/c16_pytest.txt 2>&1; tail -6 gpurun_out/c16_pytest.txt
timeout 300 python tools/gpu_diag.py --step train_fused > gpurun_out/c16_train_fused.txt 2>&1; tail -2 gpurun_out/c16_train_fused.txt
timeout 300 python tools/abtime.py toad_b200/libtoad_b200.so > gpurun_out/c16_ab.json 2>&1; cat gpurun_out/c16_ab.json
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/c16_train_launches.csv python tools/profile_train_fused.py > gpurun_out/c16_ncu_train.log 2>&1
