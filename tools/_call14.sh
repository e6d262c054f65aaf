set -x
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/c14_train_launches.csv python tools/profile_train_fused.py > gpurun_out/c14_ncu_train.log 2>&1; tail -3 gpurun_out/c14_ncu_train.log
