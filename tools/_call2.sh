set -x
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/c2_pytest.txt 2>&1; tail -15 gpurun_out/c2_pytest.txt
timeout 300 python tools/abtime.py tools/_ab/libtoad_old.so > gpurun_out/c2_ab_old.json 2>&1; cat gpurun_out/c2_ab_old.json
timeout 300 python tools/abtime.py toad_b200/libtoad_b200.so > gpurun_out/c2_ab_new.json 2>&1; cat gpurun_out/c2_ab_new.json
timeout 300 python tools/abtime.py tools/_ab/libtoad_old.so > gpurun_out/c2_ab_old2.json 2>&1; cat gpurun_out/c2_ab_old2.json
timeout 300 python tools/abtime.py toad_b200/libtoad_b200.so > gpurun_out/c2_ab_new2.json 2>&1; cat gpurun_out/c2_ab_new2.json
