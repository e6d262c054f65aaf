set -x
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16x3 -s 8 -c 8 -o gpurun_out/c19_train_gemms python tools/profile_train_fused.py > gpurun_out/c19_ncu.log 2>&1; tail -2 gpurun_out/c19_ncu.log
timeout 300 python bench.py --streams 4 --no-resnet --no-cpu-baseline --no-eager-baseline > gpurun_out/c19_bench_streams4.json 2>/dev/null; cut -c1-120 gpurun_out/c19_bench_streams4.json
