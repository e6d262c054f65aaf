set -x
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16x3 -s 3 -c 3 -o gpurun_out/c3_fwd_gemm python tools/profile_fwd.py --iters 2 > gpurun_out/c3_ncu_full.log 2>&1
