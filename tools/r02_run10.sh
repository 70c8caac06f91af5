mkdir -p gpurun_out
T=${1:-r02k}
B="--no-cpu-baseline --e2e-steps 0 --reverse-steps 0 --steps 3 --warmup 3 --dirs 1024"
for v in "def:" "cbs2_ws8:--cbs 2 --opt wide_streams=8" "cbs4_ws4_c64:--cbs 4 --opt wide_streams=4 --chunk 64" "cbs2_ws8_c64:--cbs 2 --opt wide_streams=8 --chunk 64" "cbs4_ws8:--cbs 4 --opt wide_streams=8"; do
  name=${v%%:*}; opts=${v#*:}
  timeout 900 python bench.py --workload synthetic1e8_fwd $B $opts > gpurun_out/${T}_d1024_$name.json 2> gpurun_out/${T}_d1024_$name.err
done
for f in gpurun_out/${T}_d1024_*.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1]); r=d.get('roofline') or {}
    print(sys.argv[1].split('/')[-1], 'ms', round(d['ms_per_step'],1), 'frac', round(r.get('frac',0) or 0,3), 'sweep_ms', round(r.get('sweep_ms_per_step',0),1), 'streams', r.get('concurrent_streams'), 'parity', d.get('parity_bit_exact_vs_oracle'))
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex)
PY
done
