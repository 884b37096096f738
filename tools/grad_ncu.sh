set -x
python tools/perf_probe.py --batches 256 --clusters 1 --pair 3 --grad --iters 1 > gpurun_out/grad_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_' -c 20000 --csv --log-file gpurun_out/grad_launches.csv python tools/perf_probe.py --batches 256 --clusters 1 --pair 3 --grad --iters 1 > gpurun_out/grad_ncu.log 2>&1
tail -3 gpurun_out/grad_plain.log
