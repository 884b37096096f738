# Round-2 profiling pass (run under gpurun).  Every ncu command is preceded by the same command without ncu.
set -x
O=gpurun_out
NV='--nvtx --nvtx-include roi/'
LIST="ncu --metrics gpu__time_duration.sum --clock-control none $NV --csv"
FULL="ncu --set full --clock-control none --import-source on $NV"
run() { name=$1; shift; python tools/prof_run.py "$@" > $O/r02_${name}_plain.log 2>&1; }
# ---- launch lists (per-kernel shares of one pass)
run logprob logprob --prec bf16 --n 2048 && $LIST --log-file $O/r02_launches_logprob_bf16.csv python tools/prof_run.py logprob --prec bf16 --n 2048 > $O/r02_l1.log 2>&1
run ncsn1 ncsn --version v1 --n 30 && $LIST --log-file $O/r02_launches_ncsn_v1.csv python tools/prof_run.py ncsn --version v1 --n 30 > $O/r02_l2.log 2>&1
run ncsn2 ncsn --version v2 --n 30 && $LIST --log-file $O/r02_launches_ncsn_v2.csv python tools/prof_run.py ncsn --version v2 --n 30 > $O/r02_l3.log 2>&1
ASEP_NO_GRAPH=1 python tools/prof_run.py train --n 32 > $O/r02_train_plain.log 2>&1 && ASEP_NO_GRAPH=1 $LIST --log-file $O/r02_launches_train.csv python tools/prof_run.py train --n 32 > $O/r02_l4.log 2>&1
run basis basis --prec bf16 --n 30 && $LIST --log-file $O/r02_launches_basis_bf16_n30.csv python tools/prof_run.py basis --prec bf16 --n 30 > $O/r02_l5.log 2>&1
# ---- full captures of the dominant / HBM-bound kernels
run tcx logprob --prec fp16x3 --n 512 --K 2 && $FULL -k regex:k_nn_tcx -c 2 -o $O/r02_k_nn_tcx_fp16x3 python tools/prof_run.py logprob --prec fp16x3 --n 512 --K 2 > $O/r02_f1.log 2>&1
run tc4 logprob --prec bf16 --n 512 --K 2 && $FULL -k regex:k_nn_tc4 -c 2 -o $O/r02_k_nn_tc4 python tools/prof_run.py logprob --prec bf16 --n 512 --K 2 > $O/r02_f2.log 2>&1
run pp logprob --prec bf16 --n 2048 --K 2 && $FULL -k regex:k_post_pre -c 3 -o $O/r02_k_post_pre python tools/prof_run.py logprob --prec bf16 --n 2048 --K 2 > $O/r02_f3.log 2>&1
run lan langevin --n 4096 && $FULL -k regex:k_langevin -c 2 -o $O/r02_k_langevin python tools/prof_run.py langevin --n 4096 > $O/r02_f4.log 2>&1
run prep ncsn --version v1 --n 30 && $FULL -k regex:'k_prep|k_pool5_1d' -c 6 -o $O/r02_k_prep_pool python tools/prof_run.py ncsn --version v1 --n 30 > $O/r02_f5.log 2>&1
run gemm gemm && $FULL -c 2 -o $O/r02_cublas_gemm python tools/prof_run.py gemm > $O/r02_f6.log 2>&1
tail -2 $O/r02_*_plain.log | cut -c 1-200
ls -la $O/*.ncu-rep $O/r02_launches_*.csv
