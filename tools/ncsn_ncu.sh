set -x
python tools/ncsn_probe.py --version v1 --batches 30 > gpurun_out/ncsn_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_' -c 3000 --csv --log-file gpurun_out/ncsn_launches.csv python tools/ncsn_probe.py --version v1 --batches 30 > gpurun_out/ncsn_ncu.log 2>&1
tail -2 gpurun_out/ncsn_ncu.log
