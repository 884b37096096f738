import json, sys
j = json.loads(open(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/r1_bench.json').read().strip().splitlines()[-1])
for k in ['value', 'ms_per_step', 'gpu_launches', 'n_gpus']:
    print(k, j[k])
print('e2e', j['e2e'])
print('roof', {k: j['roofline'].get(k) for k in ['achieved', 'frac', 'kernel_share_of_step', 'traffic', 'avg_launch_ms']})
print('cpu', j['cpu_baseline'])
print('clocks', j['clocks'])
if j.get('directions'):
    for k, v in j['directions'].items():
        print(k, {a: v[a] for a in v if a != 'note'})
if j.get('basis'):
    print('basis', {k: j['basis'][k] for k in ['value', 'ms_per_langevin_step', 'alg_tflops']})
for v in ['v1', 'v2']:
    if j.get('basis_ncsn') and v in j['basis_ncsn']:
        b = j['basis_ncsn'][v]
        print(v, {k: b[k] for k in ['value', 'ms_per_langevin_step', 'alg_tflops']}, b['roofline']['frac'], b['roofline']['kernel_share_of_step'])
if j.get('train'):
    print('train', {k: j['train'][k] for k in ['value', 'ms_per_step', 'alg_tflops', 'loss']})
