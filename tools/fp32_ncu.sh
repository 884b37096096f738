set -x
python tools/perf_probe.py --K 4 --batches 32 --fp32 --iters 1 > gpurun_out/fp32_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_conv1|k_sgemm|k_conv3' -c 400 --csv --log-file gpurun_out/fp32_launches.csv python tools/perf_probe.py --K 4 --batches 32 --fp32 --iters 1 > gpurun_out/fp32_ncu.log 2>&1
tail -2 gpurun_out/fp32_plain.log
