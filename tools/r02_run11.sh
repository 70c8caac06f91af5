# Round 2, GPU run 11: the GPU test suite against the bounds-checked build, then against the product build
mkdir -p gpurun_out
T=${1:-r02c2}
( time ADEPT_CUDA_LIB=$PWD/adept-2_b200/libadept_cuda_check.so timeout 2400 python -m pytest tests -m gpu -q -x ) > gpurun_out/${T}_pytest_check.log 2>&1
tail -5 gpurun_out/${T}_pytest_check.log
( time timeout 2400 python -m pytest tests -m gpu -q ) > gpurun_out/${T}_pytest.log 2>&1
tail -5 gpurun_out/${T}_pytest.log
ADEPT_CUDA_LIB=$PWD/adept-2_b200/libadept_cuda_check.so timeout 900 python bench.py --workload synthetic1e7_fwd --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/${T}_bench_check.json 2> gpurun_out/${T}_bench_check.err
python - <<'PY'
import json,os
T=os.environ.get('T','r02c2')
try:
    d=json.loads([l for l in open(f'gpurun_out/{T}_bench_check.json') if l.startswith('{')][-1])
    print('bounds-checked bench: ms', d['ms_per_step'], 'parity', d['parity_bit_exact_vs_oracle'], 'rev', d['reverse']['ms_per_step'])
except Exception as e:
    print('bench under check build FAILED', e); print(open(f'gpurun_out/{T}_bench_check.err').read()[-800:])
PY
