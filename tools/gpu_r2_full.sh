set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 400 python bench.py 2>gpurun_out/bench_err.log | tail -1 > gpurun_out/bench_n1_late.json; tail -c 600 gpurun_out/bench_n1_late.json
timeout 300 python bench.py --impl reference 2>>gpurun_out/bench_err.log | tail -1 > gpurun_out/bench_ref_late.json; cat gpurun_out/bench_ref_late.json | cut -c1-400
