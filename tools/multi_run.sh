# multi-GPU check (run under gpurun --gpus N): N-device parity tests + the default bench under torchrun
N=${1:-2}
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_dropin.py -m gpu -x -q ) > gpurun_out/multi_pytest_n$N.log 2>&1
tail -3 gpurun_out/multi_pytest_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/multi_bench_toon4096_n$N.json 2> gpurun_out/multi_bench_toon4096_n$N.err
tail -2 gpurun_out/multi_bench_toon4096_n$N.err
python - gpurun_out/multi_bench_toon4096_n$N.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1]); e=d.get('e2e') or {}
    print('n_gpus', d['n_gpus'], 'ms', round(d['ms_per_step'],1), 'value %.3e' % d['value'], 'e2e %.3e' % (e.get('value') or 0), 'parity', d.get('parity_bit_exact_vs_oracle'))
except Exception as ex:
    print('FAILED', ex)
PY
