set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r1_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1_smoke.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err
tail -3 gpurun_out/r1_tests.log; cat gpurun_out/r1_smoke.log | tail -3; cat gpurun_out/r1_bench.json; tail -5 gpurun_out/r1_bench.err
