set -x
python tools/train_probe.py --K 4 --steps 3 --lr 2e-5 > gpurun_out/train_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_|vectorized|elementwise' -c 4000 --csv --log-file gpurun_out/train_launches.csv python tools/train_probe.py --K 4 --steps 3 --lr 2e-5 > gpurun_out/train_ncu.log 2>&1
tail -2 gpurun_out/train_ncu.log
