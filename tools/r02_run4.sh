# Round 2, GPU run 4: tests of the new pieces, forward-kernel tuning sweep, config 5 / config 3 lines, analysis probe
mkdir -p gpurun_out
T=${1:-r02d}
( time timeout 2400 python -m pytest tests -m gpu -q ) > gpurun_out/${T}_pytest.log 2>&1
tail -8 gpurun_out/${T}_pytest.log
B="--no-cpu-baseline --e2e-steps 0 --reverse-steps 0 --steps 3 --warmup 3"
for v in "cbs8:--cbs 8" "cbs16:--cbs 16" "cbs32:--cbs 32" "cbs16_c64:--cbs 16 --chunk 64" "cbs16_c192:--cbs 16 --chunk 192" "cbs16_c224:--cbs 16 --chunk 224" "cbs32_c192:--cbs 32 --chunk 192" "cbs16_ws2:--cbs 16 --opt wide_streams=2" "cbs8_ws2:--cbs 8 --opt wide_streams=2"; do
  name=${v%%:*}; opts=${v#*:}
  timeout 600 python bench.py --workload synthetic1e7_fwd $B $opts > gpurun_out/${T}_syn1e7_$name.json 2> gpurun_out/${T}_syn1e7_$name.err
done
python tools/analysis_probe.py synthetic1e7_fwd 2 > gpurun_out/${T}_probe_syn1e7.out 2>&1; tail -1 gpurun_out/${T}_probe_syn1e7.out
python tools/analysis_probe.py rosenbrock1e7 2 > gpurun_out/${T}_probe_rosen.out 2>&1; tail -1 gpurun_out/${T}_probe_rosen.out
timeout 900 python bench.py --workload rosenbrock1e7 --steps 5 > gpurun_out/${T}_bench_rosenbrock1e7.json 2> gpurun_out/${T}_bench_rosenbrock1e7.err
timeout 900 python bench.py --workload ensemble1e5 --steps 3 > gpurun_out/${T}_bench_ensemble1e5.json 2> gpurun_out/${T}_bench_ensemble1e5.err
cut -c1-1500 gpurun_out/${T}_bench_rosenbrock1e7.json; cut -c1-1800 gpurun_out/${T}_bench_ensemble1e5.json; tail -3 gpurun_out/${T}_bench_ensemble1e5.err
timeout 600 python bench.py --workload toon4096 --no-cpu-baseline > gpurun_out/${T}_bench_toon4096.json 2> gpurun_out/${T}_bench_toon4096.err
for f in gpurun_out/${T}_syn1e7_*.json gpurun_out/${T}_bench_toon4096.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1]); r=d.get('roofline') or {}; e=d.get('e2e') or {}
    print(sys.argv[1].split('/')[-1], 'ms', round(d['ms_per_step'],1), 'frac', round(r.get('frac',0) or 0,3), 'sweep_ms', round(r.get('sweep_ms_per_step',0),1), 'streams', r.get('concurrent_streams'), 'parity', d.get('parity_bit_exact_vs_oracle'), 'e2e_ms', e.get('ms_per_step'))
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex)
PY
done
