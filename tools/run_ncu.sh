set -x
ARGS="--batch 512 --steps 1 --warmup 3 --basis-segments 0 --cpu-sample 0 --ncsn-segments 0 --train-batch 0"
python bench.py $ARGS > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_' -s 1107 -c 369 --csv --log-file gpurun_out/r01_launches.csv python bench.py $ARGS > gpurun_out/ncu_list.log 2>&1
python bench.py $ARGS > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_nn_tc -s 360 -c 2 -o gpurun_out/r01_k_nn_tc python bench.py $ARGS > gpurun_out/ncu_full.log 2>&1
python tools/ncsn_probe.py --version v1 --batches 30 > gpurun_out/ncsn_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 250 -c 2 -o gpurun_out/r01_k_conv_tc python tools/ncsn_probe.py --version v1 --batches 30 > gpurun_out/ncu_full_conv.log 2>&1
tail -c 300 gpurun_out/ncu_plain.log; tail -2 gpurun_out/ncu_list.log | cut -c 1-200; tail -2 gpurun_out/ncu_full.log | cut -c 1-200; tail -2 gpurun_out/ncu_full_conv.log | cut -c 1-200; ls -la gpurun_out/*.ncu-rep gpurun_out/r01_launches.csv
