set -x
(time timeout 900 python -m pytest tests/test_gpu_resnet.py tests/test_gpu_forward.py -m gpu -x -q) > gpurun_out/c11_pytest.txt 2>&1; tail -5 gpurun_out/c11_pytest.txt
timeout 600 python tools/resnet_sweep.py 64 128 256 > gpurun_out/c11_resnet_c32.json 2>&1; cat gpurun_out/c11_resnet_c32.json
TOAD_B200_LIB=tools/_ab/libtoad_rc16.so timeout 600 python tools/resnet_sweep.py 128 256 > gpurun_out/c11_resnet_c16.json 2>&1; cat gpurun_out/c11_resnet_c16.json
TOAD_B200_LIB=tools/_ab/libtoad_rc64.so timeout 600 python tools/resnet_sweep.py 128 256 > gpurun_out/c11_resnet_c64.json 2>&1; cat gpurun_out/c11_resnet_c64.json
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/c11_resnet_launches.csv python tools/profile_resnet.py --batch 128 --iters 2 > gpurun_out/c11_ncu_resnet.log 2>&1
