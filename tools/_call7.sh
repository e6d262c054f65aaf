set -x
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/c8_pytest.txt 2>&1; tail -15 gpurun_out/c8_pytest.txt
timeout 300 python tools/abtime.py toad_b200/libtoad_b200.so > gpurun_out/c8_ab.json 2>&1; cat gpurun_out/c8_ab.json
for k in 1 2 3 4 5; do TOAD_TAIL_STOP=$k timeout 300 python tools/abtime.py tools/_ab/libtoad_taildbg.so > gpurun_out/c8_tail_stop$k.json 2>&1; cat gpurun_out/c8_tail_stop$k.json; done
timeout 300 python tools/abtime.py toad_b200/libtoad_b200.so 0 10000 > gpurun_out/c8_ab10k.json 2>&1; cat gpurun_out/c8_ab10k.json
timeout 300 python tools/abtime.py toad_b200/libtoad_b200.so 0 200000 > gpurun_out/c8_ab200k.json 2>&1; cat gpurun_out/c8_ab200k.json
