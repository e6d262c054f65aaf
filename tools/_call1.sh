set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r1m_pytest.txt 2>&1; tail -5 gpurun_out/r1m_pytest.txt
timeout 120 python __graft_entry__.py --smoke > gpurun_out/r1m_smoke.txt 2>&1; tail -2 gpurun_out/r1m_smoke.txt
timeout 600 python bench.py > gpurun_out/r1m_bench.json 2> gpurun_out/r1m_bench.err; tail -c 600 gpurun_out/r1m_bench.json
timeout 300 python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/r1m_bench_reference.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1m_launches.csv python bench.py --steps 2 --warmup 3 --slides-per-step 4 --no-cpu-baseline --no-resnet > gpurun_out/r1m_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16x3 -s 3 -c 3 -o gpurun_out/r1m_fwd_gemm python tools/profile_fwd.py --iters 2 > gpurun_out/r1m_ncu_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pool_heads -s 1 -c 1 -o gpurun_out/r1m_fwd_tail python tools/profile_fwd.py --iters 2 >> gpurun_out/r1m_ncu_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r1m_resnet_launches.csv python tools/profile_resnet.py --batch 128 --iters 2 > gpurun_out/r1m_ncu_resnet.log 2>&1
ls -la gpurun_out
