set -x
timeout 300 python tools/abtime.py toad_b200/libtoad_b200.so > gpurun_out/c6_ab_v0.json 2>&1; cat gpurun_out/c6_ab_v0.json
timeout 300 python tools/abtime.py tools/_ab/libtoad_epi8.so > gpurun_out/c6_ab_epi8.json 2>&1; cat gpurun_out/c6_ab_epi8.json
for k in 0 1 2 3 4 5; do TOAD_TAIL_STOP=$k timeout 300 python tools/abtime.py tools/_ab/libtoad_taildbg.so > gpurun_out/c6_tail_stop$k.json 2>&1; cat gpurun_out/c6_tail_stop$k.json; done
for k in 0 1 2 3; do TOAD_TAIL_STOP=$k timeout 300 python tools/abtime.py tools/_ab/libtoad_taildbg.so 0 10000 > gpurun_out/c6_tail10k_stop$k.json 2>&1; cat gpurun_out/c6_tail10k_stop$k.json; done
