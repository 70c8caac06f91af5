# Round 2, GPU run 6: prefetching window walk -- tests, probes, config 5
mkdir -p gpurun_out
T=${1:-r02f}
( time timeout 2400 python -m pytest tests -m gpu -q ) > gpurun_out/${T}_pytest.log 2>&1
tail -6 gpurun_out/${T}_pytest.log
python tools/analysis_probe.py synthetic1e7_fwd 2 > gpurun_out/${T}_probe_syn1e7.out 2>&1; tail -1 gpurun_out/${T}_probe_syn1e7.out
python tools/analysis_probe.py rosenbrock1e7 2 > gpurun_out/${T}_probe_rosen.out 2>&1; tail -1 gpurun_out/${T}_probe_rosen.out
python tools/analysis_probe.py synthetic1e8_fwd 2 > gpurun_out/${T}_probe_syn1e8.out 2>&1; tail -1 gpurun_out/${T}_probe_syn1e8.out
timeout 900 python bench.py --workload rosenbrock1e7 --steps 5 > gpurun_out/${T}_bench_rosenbrock1e7.json 2> gpurun_out/${T}_bench_rosenbrock1e7.err
python - <<'PY'
import json,os
T=os.environ.get('T','r02f')
for f in (f"gpurun_out/{T}_bench_rosenbrock1e7.json",):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1]); e=d.get('e2e') or {}
        print(f.split('/')[-1], 'value %.3e' % d['value'], 'ms', round(d['ms_per_step'],2), 'e2e_ms', e.get('ms_per_step'), 'ref', e.get('reference_sweep_ms'), 'parity', d.get('parity_bit_exact_vs_oracle'))
        print('   ', e.get('phases_ms'))
    except Exception as ex:
        print(f, 'FAILED', ex)
PY
