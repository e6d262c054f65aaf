set -x
(time timeout 600 python -m pytest tests -m gpu -x -q) > gpurun_out/r1o_pytest.txt 2>&1; tail -4 gpurun_out/r1o_pytest.txt
timeout 120 python __graft_entry__.py --smoke > gpurun_out/r1o_smoke.txt 2>&1; tail -1 gpurun_out/r1o_smoke.txt
timeout 900 python bench.py > gpurun_out/r1o_bench.json 2> gpurun_out/r1o_bench.err; tail -c 300 gpurun_out/r1o_bench.err; cut -c1-200 gpurun_out/r1o_bench.json
