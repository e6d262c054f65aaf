# Round-end evidence run on the GPU box (under gpurun): bench, reference arm, ncu launch lists and full captures,
# compute-sanitizer.  TAG=r2z bash tools/gpu_round_end.sh ; then summarise with tools/summarize_ncu.py /
# tools/resnet_launch_summary.py into profiles/.
set -x
T=${TAG:-rXX}
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 400 gpurun_out/${T}_bench.err
timeout 300 python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/${T}_bench_reference.json 2>/dev/null
# launch list of the SAME command as the bench's headline region (short: ncu serialises and replays)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 1 --warmup 3 --slides-per-step 4 --no-cpu-baseline --no-resnet --no-eager-baseline --no-train --no-traffic > gpurun_out/${T}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16x3|pool_heads" -s 4 -c 4 -o gpurun_out/${T}_fwd python tools/profile_fwd.py --iters 2 > gpurun_out/${T}_ncu_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/${T}_resnet_launches.csv python tools/profile_resnet.py --batch 128 --iters 2 > gpurun_out/${T}_ncu_resnet.log 2>&1
python tools/resnet_launch_summary.py gpurun_out/${T}_resnet_launches.csv list > gpurun_out/${T}_resnet_launches.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_train_launches.csv python tools/profile_train_fused.py > gpurun_out/${T}_ncu_train.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_small.py > gpurun_out/${T}_sanitizer_memcheck.txt 2>&1; echo "sanitizer rc=$?"; tail -4 gpurun_out/${T}_sanitizer_memcheck.txt
