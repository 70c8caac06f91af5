# Round 2, GPU run 3: balanced forward kernel -- tests + bench comparisons
mkdir -p gpurun_out
T=${1:-r02c}
( time timeout 2400 python -m pytest tests -m gpu -q ) > gpurun_out/${T}_pytest.log 2>&1
tail -8 gpurun_out/${T}_pytest.log
B="--no-cpu-baseline --e2e-steps 0 --reverse-steps 0 --steps 3 --warmup 3"
timeout 600 python bench.py --workload synthetic1e7_fwd $B --opt fwd_balanced=0 > gpurun_out/${T}_syn1e7_old.json 2> gpurun_out/${T}_syn1e7_old.err
timeout 600 python bench.py --workload synthetic1e7_fwd $B > gpurun_out/${T}_syn1e7_new.json 2> gpurun_out/${T}_syn1e7_new.err
timeout 600 python bench.py --workload synthetic1e7_fwd $B --opt wide_streams=2 > gpurun_out/${T}_syn1e7_new_ws2.json 2> gpurun_out/${T}_syn1e7_new_ws2.err
timeout 600 python bench.py --workload synthetic1e7_fwd $B --cbs 16 > gpurun_out/${T}_syn1e7_new_cbs16.json 2> gpurun_out/${T}_syn1e7_new_cbs16.err
timeout 900 python bench.py $B > gpurun_out/${T}_cfg4_new.json 2> gpurun_out/${T}_cfg4_new.err
for f in gpurun_out/${T}_*.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1]); r=d.get('roofline') or {}
    print(sys.argv[1].split('/')[-1], 'ms', round(d['ms_per_step'],1), 'frac', round(r.get('frac',0) or 0,3), 'sweep_ms', round(r.get('sweep_ms_per_step',0),1), 'streams', r.get('concurrent_streams'), 'parity', d.get('parity_bit_exact_vs_oracle'))
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex)
PY
done
