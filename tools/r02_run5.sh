# Round 2, GPU run 5: four-statement window walk -- tests, analysis probes, config 5 line, default bench (config 4) with e2e
mkdir -p gpurun_out
T=${1:-r02e}
( time timeout 2400 python -m pytest tests -m gpu -q ) > gpurun_out/${T}_pytest.log 2>&1
tail -6 gpurun_out/${T}_pytest.log
python tools/analysis_probe.py synthetic1e7_fwd 2 > gpurun_out/${T}_probe_syn1e7.out 2>&1; tail -1 gpurun_out/${T}_probe_syn1e7.out
python tools/analysis_probe.py rosenbrock1e7 2 > gpurun_out/${T}_probe_rosen.out 2>&1; tail -1 gpurun_out/${T}_probe_rosen.out
python tools/analysis_probe.py toon4096 2 > gpurun_out/${T}_probe_toon.out 2>&1; tail -1 gpurun_out/${T}_probe_toon.out
timeout 900 python bench.py --workload rosenbrock1e7 --steps 5 > gpurun_out/${T}_bench_rosenbrock1e7.json 2> gpurun_out/${T}_bench_rosenbrock1e7.err
timeout 900 python bench.py --workload checkpoint --steps 2 > gpurun_out/${T}_bench_checkpoint.json 2> gpurun_out/${T}_bench_checkpoint.err
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench_cfg4.json 2> gpurun_out/${T}_bench_cfg4.err
python - <<'PY'
import json,os
T=os.environ.get('T','r02e')
for f in (f"gpurun_out/{T}_bench_rosenbrock1e7.json", f"gpurun_out/{T}_bench_checkpoint.json", f"gpurun_out/{T}_bench_cfg4.json"):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1]); e=d.get('e2e') or {}; r=d.get('roofline') or {}; o=d.get('reverse') or {}
        print(f.split('/')[-1], 'value %.3e' % d['value'], 'ms', round(d['ms_per_step'],2), 'frac', r.get('frac'), 'e2e_ms', e.get('ms_per_step'), 'parity', d.get('parity_bit_exact_vs_oracle'), 'rev', o.get('ms_per_step'), (o.get('roofline') or {}).get('frac'))
        print('   ', e.get('phases_ms')); print('   ', e.get('verdict')); print('   cpu', d.get('cpu_baseline'))
    except Exception as ex:
        print(f, 'FAILED', ex)
PY
