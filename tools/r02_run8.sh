# Round 2, GPU run 8: CAS-free walk (tests + probes), CTA-shape rule at 1024 / 2048 / 4096 / 8192 directions
mkdir -p gpurun_out
T=${1:-r02i}
( time timeout 2400 python -m pytest tests -m gpu -q ) > gpurun_out/${T}_pytest.log 2>&1
tail -4 gpurun_out/${T}_pytest.log
python tools/analysis_probe.py synthetic1e7_fwd 2 > gpurun_out/${T}_probe_syn1e7.out 2>&1; tail -1 gpurun_out/${T}_probe_syn1e7.out
python tools/analysis_probe.py rosenbrock1e7 2 > gpurun_out/${T}_probe_rosen.out 2>&1; tail -1 gpurun_out/${T}_probe_rosen.out
B="--no-cpu-baseline --e2e-steps 0 --reverse-steps 0 --steps 3 --warmup 3"
for v in "d1024_def:--dirs 1024" "d2048_def:--dirs 2048" "d2048_cbs8_ws2:--dirs 2048 --cbs 8 --opt wide_streams=2" "d2048_cbs4_ws4:--dirs 2048 --cbs 4 --opt wide_streams=4" "d4096_def:--dirs 4096" "d4096_ws4:--dirs 4096 --opt wide_streams=4" "d4096_cbs8_ws4:--dirs 4096 --cbs 8 --opt wide_streams=4" "d8192_ws4:--opt wide_streams=4" "d8192_cbs8_ws4:--cbs 8 --opt wide_streams=4"; do
  name=${v%%:*}; opts=${v#*:}
  timeout 900 python bench.py --workload synthetic1e8_fwd $B $opts > gpurun_out/${T}_$name.json 2> gpurun_out/${T}_$name.err
done
timeout 900 python bench.py --workload rosenbrock1e7 --steps 5 > gpurun_out/${T}_bench_rosenbrock1e7.json 2> gpurun_out/${T}_bench_rosenbrock1e7.err
for f in gpurun_out/${T}_d*.json gpurun_out/${T}_bench_rosenbrock1e7.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1]); r=d.get('roofline') or {}; e=d.get('e2e') or {}
    print(sys.argv[1].split('/')[-1], 'ms', round(d['ms_per_step'],2), 'frac', round(r.get('frac',0) or 0,3), 'sweep_ms', round(r.get('sweep_ms_per_step',0) or 0,1), 'streams', r.get('concurrent_streams'), 'parity', d.get('parity_bit_exact_vs_oracle'), 'e2e_ms', e.get('ms_per_step'), e.get('reference_sweep_ms'))
    if e.get('phases_ms'): print('   ', {k:v for k,v in e['phases_ms'].items() if k!='each_step_upload_jacobian_replay_ms'})
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex)
PY
done
