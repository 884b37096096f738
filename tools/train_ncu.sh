set -x
export ASEP_NO_GRAPH=1
python tools/train_probe.py --K 40 --steps 2 > gpurun_out/train_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_' -c 20000 --csv --log-file gpurun_out/train_launches.csv python tools/train_probe.py --K 40 --steps 2 > gpurun_out/train_ncu.log 2>&1
tail -2 gpurun_out/train_ncu.log
