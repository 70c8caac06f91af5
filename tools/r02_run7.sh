# Round 2, GPU run 7: what one rank of an 8-GPU run sees (1024 of config 4's 8192 directions) -- CTA-shape variants;
# plus the tests of the overlapped upload / cooperative packing / pool retention and the default bench with e2e.
mkdir -p gpurun_out
T=${1:-r02h}
( time timeout 2400 python -m pytest tests -m gpu -q ) > gpurun_out/${T}_pytest.log 2>&1
tail -4 gpurun_out/${T}_pytest.log
B="--no-cpu-baseline --e2e-steps 0 --reverse-steps 0 --steps 3 --warmup 3 --dirs 1024"
for v in "def:" "cbs16:--cbs 16" "cbs4:--cbs 4" "cbs8_s1:--cbs 8 --opt wide_streams=1" "cbs16_s1:--cbs 16 --opt wide_streams=1" "cbs8_c64:--cbs 8 --chunk 64" "cbs4_c64:--cbs 4 --chunk 64" "cbs8_s4:--cbs 4 --opt wide_streams=4" "old:--opt fwd_balanced=0"; do
  name=${v%%:*}; opts=${v#*:}
  timeout 900 python bench.py --workload synthetic1e8_fwd $B $opts > gpurun_out/${T}_d1024_$name.json 2> gpurun_out/${T}_d1024_$name.err
done
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench_cfg4.json 2> gpurun_out/${T}_bench_cfg4.err
for f in gpurun_out/${T}_d1024_*.json gpurun_out/${T}_bench_cfg4.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1]); r=d.get('roofline') or {}; e=d.get('e2e') or {}
    print(sys.argv[1].split('/')[-1], 'ms', round(d['ms_per_step'],1), 'frac', round(r.get('frac',0) or 0,3), 'sweep_ms', round(r.get('sweep_ms_per_step',0),1), 'streams', r.get('concurrent_streams'), 'parity', d.get('parity_bit_exact_vs_oracle'), 'e2e_ms', e.get('ms_per_step'))
    if e.get('phases_ms'): print('   ', {k:v for k,v in e['phases_ms'].items() if k!='each_step_upload_jacobian_replay_ms'})
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex)
PY
done
