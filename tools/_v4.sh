mkdir -p gpurun_out
T=r02v
( time timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "four_doubles or fwd_v4 or forward_level_sweep_both or golden_jacobians" ) > gpurun_out/${T}_pytest.log 2>&1
tail -4 gpurun_out/${T}_pytest.log
B="--no-cpu-baseline --e2e-steps 0 --reverse-steps 0 --steps 3 --warmup 3"
for v in "v2:" "v4:--opt fwd_v4=1" "v4_cbs4:--opt fwd_v4=1 --cbs 4" "v4_cbs16:--opt fwd_v4=1 --cbs 16"; do
  name=${v%%:*}; opts=${v#*:}
  timeout 900 python bench.py --workload synthetic1e8_fwd $B $opts > gpurun_out/${T}_$name.json 2> gpurun_out/${T}_$name.err
  python - gpurun_out/${T}_$name.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1]); r=d.get('roofline') or {}
    print(sys.argv[1].split('/')[-1], 'ms', round(d['ms_per_step'],1), 'frac', round(r.get('frac',0) or 0,3), 'sweep_ms', round(r.get('sweep_ms_per_step',0),1), 'streams', r.get('concurrent_streams'), 'parity', d.get('parity_bit_exact_vs_oracle'), d['clocks'])
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex); print(open(sys.argv[1].replace('.json','.err')).read()[-500:])
PY
done
