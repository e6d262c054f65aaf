set -x
timeout 600 python tools/resnet_sweep.py > gpurun_out/c10_resnet_sweep.json 2>&1; cat gpurun_out/c10_resnet_sweep.json
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/c10_resnet_launches.csv python tools/profile_resnet.py --batch 128 --iters 2 > gpurun_out/c10_ncu_resnet.log 2>&1
