set -x
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/c5_pytest.txt 2>&1; tail -15 gpurun_out/c5_pytest.txt
timeout 300 python tools/abtime.py toad_b200/libtoad_b200.so > gpurun_out/c5_ab_new.json 2>&1; cat gpurun_out/c5_ab_new.json
timeout 300 python tools/abtime.py toad_b200/libtoad_b200.so 0 200000 > gpurun_out/c5_ab_new_200k.json 2>&1; cat gpurun_out/c5_ab_new_200k.json
timeout 300 python tools/abtime.py toad_b200/libtoad_b200.so 0 10000 > gpurun_out/c5_ab_new_10k.json 2>&1; cat gpurun_out/c5_ab_new_10k.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pool_heads" -s 1 -c 1 -o gpurun_out/c5_tail python tools/profile_fwd.py --iters 2 > gpurun_out/c5_ncu_full.log 2>&1
