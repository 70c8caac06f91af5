# Round 2 multi-GPU run (under gpurun --gpus N): multi-device / one-process-per-GPU parity tests, the default bench
# (BASELINE config 4, full size) under torchrun, config 2 under torchrun, config 3 sharded over N devices.
N=${1:-2}
T=${2:-r02m}
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_multi.py tests/test_gpu_torchrun.py tests/test_ensemble_helper.py -m gpu -q -rs ) > gpurun_out/${T}_pytest_n$N.log 2>&1
tail -12 gpurun_out/${T}_pytest_n$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 1500 $TR --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_cfg4_n$N.json 2> gpurun_out/${T}_bench_cfg4_n$N.err
tail -2 gpurun_out/${T}_bench_cfg4_n$N.err
timeout 600 $TR --master-port 29518 bench.py --gpus $N --workload toon4096 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_toon4096_n$N.json 2> gpurun_out/${T}_bench_toon4096_n$N.err
timeout 600 python bench.py --gpus $N --workload ensemble1e5 --steps 3 > gpurun_out/${T}_bench_ensemble1e5_n$N.json 2> gpurun_out/${T}_bench_ensemble1e5_n$N.err
for f in gpurun_out/${T}_bench_*_n$N.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1]); e=d.get('e2e') or {}; r=d.get('roofline') or {}; o=d.get('reverse') or {}
    print(sys.argv[1].split('/')[-1], 'n_gpus', d.get('n_gpus'), 'ms', round(d['ms_per_step'],1), 'value %.3e' % d['value'], 'frac', r.get('frac'), 'e2e_ms', e.get('ms_per_step'),
          'parity', d.get('parity_bit_exact_vs_oracle', d.get('parity_bit_exact_vs_reference_every_member')), 'rev_ms', o.get('ms_per_step'))
    print('   phases', e.get('phases_ms'))
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex)
PY
done
