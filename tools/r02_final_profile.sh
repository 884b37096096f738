# Final round-2 measurement pass (run under gpurun).  Every ncu command is preceded by the same command without ncu.
set -x
O=gpurun_out
NV='--nvtx --nvtx-include roi/'
LIST="ncu --metrics gpu__time_duration.sum --clock-control none $NV --csv"
FULL="ncu --set full --clock-control none --import-source on $NV"
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_final_smoke.log 2>&1; tail -2 $O/r02_final_smoke.log
python bench.py > $O/r02_bench_final.json 2> $O/r02_bench_final.err; tail -c 300 $O/r02_bench_final.json
python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err; tail -c 300 $O/r02_bench_reference.json
run() { name=$1; shift; python tools/prof_run.py "$@" > $O/r02f_${name}_plain.log 2>&1; }
run nt1 ncsn_train --version v1 --n 32 && $LIST --log-file $O/r02_launches_ncsn_train_v1.csv python tools/prof_run.py ncsn_train --version v1 --n 32 > $O/r02f_l1.log 2>&1
run nt2 ncsn_train --version v2 --n 32 && $LIST --log-file $O/r02_launches_ncsn_train_v2.csv python tools/prof_run.py ncsn_train --version v2 --n 32 > $O/r02f_l2.log 2>&1
run n1 ncsn --version v1 --n 30 && $LIST --log-file $O/r02_launches_ncsn_v1.csv python tools/prof_run.py ncsn --version v1 --n 30 > $O/r02f_l3.log 2>&1
run n2 ncsn --version v2 --n 30 && $LIST --log-file $O/r02_launches_ncsn_v2.csv python tools/prof_run.py ncsn --version v2 --n 30 > $O/r02f_l4.log 2>&1
run wg ncsn_train --version v1 --n 32 && $FULL -k regex:k_conv_wgrad_tc -c 3 -o $O/r02_k_conv_wgrad_tc python tools/prof_run.py ncsn_train --version v1 --n 32 > $O/r02f_f1.log 2>&1
run pb ncsn_train --version v1 --n 32 && $FULL -k regex:'k_prep_bwd_apply|k_prep_bwd_reduce' -c 4 -o $O/r02_k_prep_bwd python tools/prof_run.py ncsn_train --version v1 --n 32 > $O/r02f_f2.log 2>&1
tail -2 $O/r02f_*_plain.log | cut -c 1-200
ls -la $O/*.ncu-rep $O/r02_launches_*.csv | tail -20
