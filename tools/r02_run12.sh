# Round 2, GPU run 12: fused reverse sweep on config 4's shape (1e7 statements): CTA width and stream count
mkdir -p gpurun_out
T=${1:-r02r}
B="--workload synthetic1e7_rev --no-cpu-baseline --e2e-steps 0 --steps 3 --warmup 3"
for v in "def:" "cbs8:--cbs 8" "cbs32:--cbs 32" "cbs16_s8:--cbs 16 --streams 8" "cbs8_s8:--cbs 8 --streams 8" "cbs32_s2:--cbs 32 --streams 2" "cbs16_s2:--cbs 16 --streams 2"; do
  name=${v%%:*}; opts=${v#*:}
  timeout 600 python bench.py $B $opts > gpurun_out/${T}_rev_$name.json 2> gpurun_out/${T}_rev_$name.err
done
for f in gpurun_out/${T}_rev_*.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1]); r=d.get('roofline') or {}
    print(sys.argv[1].split('/')[-1], 'ms', round(d['ms_per_step'],1), 'frac', round(r.get('frac',0) or 0,3), 'sweep_ms', round(r.get('sweep_ms_per_step',0),1), 'streams', r.get('concurrent_streams'), 'launches', r.get('launches_per_step'), 'parity', d.get('parity_bit_exact_vs_oracle'))
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex)
PY
done
