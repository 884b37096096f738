set -x
export MASTER_ADDR=127.0.0.1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
ARGS="u1 u2 --model_type glow --synthetic --random_init 7 --n_mixed 5 --T 2 --K 2 --L 3 --n_filters 512 --learntop --sigma1 0.05 --sigmaL 0.01 --num_classes 2 --progression logarithmic --seed 5"
python run_basis_sep.py $ARGS --output gpurun_out/mg_sep1 > gpurun_out/mg_sep1.log 2>&1
$TR run_basis_sep.py $ARGS --output gpurun_out/mg_sep2 > gpurun_out/mg_sep2.log 2>&1
python tools/dp_train_check.py --out gpurun_out/mg_t1.npy > gpurun_out/mg_t1.log 2>&1
$TR tools/dp_train_check.py --out gpurun_out/mg_t2.npy > gpurun_out/mg_t2.log 2>&1
python - <<'PY'
import numpy as np
a, b = np.load("gpurun_out/mg_sep1/results.npz"), np.load("gpurun_out/mg_sep2/results.npz")
print("BASIS 1 vs 2 GPUs identical:", all(np.array_equal(a[k], b[k]) for k in ("x1", "x2", "mixed")),
      "max|d|", max(float(np.abs(a[k] - b[k]).max()) for k in ("x1", "x2")))
c, d = np.load("gpurun_out/mg_sep1/results_convergence.npz"), np.load("gpurun_out/mg_sep2/results_convergence.npz")
print("convergence identical:", np.array_equal(c["x1"], d["x1"]) and np.array_equal(c["x2"], d["x2"]))
t1, t2 = np.load("gpurun_out/mg_t1.npy"), np.load("gpurun_out/mg_t2.npy")
print("DP train 1 vs 2 GPUs: max|dtheta|", float(np.abs(t1[:-2] - t2[:-2]).max()), "losses", t1[-2:], t2[-2:])
PY
$TR bench.py --gpus 2 --steps 3 --warmup 3 --cpu-sample 0 > gpurun_out/mg_bench2.json 2> gpurun_out/mg_bench2.err
tail -2 gpurun_out/mg_sep2.log gpurun_out/mg_t2.log; cat gpurun_out/mg_bench2.json | cut -c 1-600; tail -3 gpurun_out/mg_bench2.err
