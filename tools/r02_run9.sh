# Round 2, GPU run 9: full ncu capture of a WIDE forward step; config 2 and config 5 with the clock sampler started before the warm-up
mkdir -p gpurun_out
T=${1:-r02z}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:forward_balanced -s 1254 -c 1 -o gpurun_out/${T}_ncu_forward_balanced -f \
   python bench.py --workload synthetic1e7_fwd --no-cpu-baseline --e2e-steps 0 --reverse-steps 0 --steps 1 --warmup 3 --opt wide_streams=1 > gpurun_out/${T}_ncu_full.log 2>&1
tail -2 gpurun_out/${T}_ncu_full.log
timeout 600 python bench.py --workload toon4096 > gpurun_out/${T}_bench_toon4096.json 2> gpurun_out/${T}_bench_toon4096.err
timeout 600 python bench.py --workload toon4096 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_toon4096_b.json 2> gpurun_out/${T}_bench_toon4096_b.err
timeout 900 python bench.py --workload rosenbrock1e7 --steps 5 > gpurun_out/${T}_bench_rosenbrock1e7.json 2> gpurun_out/${T}_bench_rosenbrock1e7.err
python __graft_entry__.py smoke 2>&1 | tail -2
for f in gpurun_out/${T}_bench_toon4096.json gpurun_out/${T}_bench_toon4096_b.json gpurun_out/${T}_bench_rosenbrock1e7.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1]); r=d.get('roofline') or {}; e=d.get('e2e') or {}
    print(sys.argv[1].split('/')[-1], 'ms', round(d['ms_per_step'],2), 'sweep_ms', r.get('sweep_ms_per_step'), 'e2e_ms', e.get('ms_per_step'), e.get('reference_sweep_ms'), d.get('clocks'), 'parity', d.get('parity_bit_exact_vs_oracle'))
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex)
PY
done
