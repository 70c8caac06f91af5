mkdir -p gpurun_out
T=r02p
( time timeout 600 python -m pytest tests/test_gpu_resident.py -m gpu -q -x ) > gpurun_out/${T}_pytest.log 2>&1
tail -6 gpurun_out/${T}_pytest.log
B="--workload toon4096 --no-cpu-baseline --e2e-steps 0 --steps 10 --warmup 3"
for v in "def:" "res:--opt fused_persistent=1" "res_cbs16:--opt fused_persistent=1 --cbs 16" "res_cbs8:--opt fused_persistent=1 --cbs 8" \
         "d512_def:--dirs 512" "d512_res:--dirs 512 --opt fused_persistent=1" "d512_res_k32:--dirs 512 --opt fused_persistent=1 --chunk 32" \
         "res_k32:--opt fused_persistent=1 --chunk 32"; do
  name=${v%%:*}; opts=${v#*:}
  timeout 300 python bench.py $B $opts > gpurun_out/${T}_$name.json 2> gpurun_out/${T}_$name.err
  python - gpurun_out/${T}_$name.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1]); r=d.get('roofline') or {}
    print(sys.argv[1].split('/')[-1], 'ms', round(d['ms_per_step'],2), 'sweep_ms', round(r.get('sweep_ms_per_step',0),2), 'launches', r.get('launches_per_step'), 'parity', d.get('parity_bit_exact_vs_oracle'), d['clocks'].get('sm_mhz'))
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex); print(open(sys.argv[1].replace('.json','.err')).read()[-600:])
PY
done
