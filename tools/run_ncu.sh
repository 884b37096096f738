set -x
ARGS="--batch 512 --steps 1 --warmup 3 --basis-segments 0 --cpu-sample 0 --ncsn-segments 0 --train-batch 0"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1_smoke.log 2>&1; tail -2 gpurun_out/r1_smoke.log
python bench.py $ARGS > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_' -s 1107 -c 369 --csv --log-file gpurun_out/r01_launches.csv python bench.py $ARGS > gpurun_out/ncu_list.log 2>&1
python bench.py $ARGS > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_nn_tc -s 360 -c 2 -o gpurun_out/r01_k_nn_tc python bench.py $ARGS > gpurun_out/ncu_full.log 2>&1
tail -c 600 gpurun_out/ncu_plain.log; tail -3 gpurun_out/ncu_list.log | cut -c 1-300; tail -3 gpurun_out/ncu_full.log | cut -c 1-300; ls -la gpurun_out
