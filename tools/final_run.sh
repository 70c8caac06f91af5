# Round-end measurement script (run under gpurun): tests, bench lines, ncu launch list and full captures.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/final_pytest.log 2>&1
tail -3 gpurun_out/final_pytest.log
timeout 400 python bench.py > gpurun_out/final_bench_toon4096.json 2> gpurun_out/final_bench_toon4096.err
timeout 400 python bench.py --impl reference > gpurun_out/final_bench_toon4096_ref.json 2> gpurun_out/final_bench_toon4096_ref.err
timeout 400 python bench.py --workload synthetic1e7_rev > gpurun_out/final_bench_synthetic1e7_rev.json 2> gpurun_out/final_bench_synthetic1e7_rev.err
timeout 300 python bench.py --workload toon1024 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/final_bench_toon1024.json 2>/dev/null
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1200 --csv --log-file gpurun_out/final_launches_toon1024.csv python bench.py --workload toon1024 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/final_ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:reverse_fused -s 2803 -c 2 -o gpurun_out/final_ncu_toon4096 -f python bench.py --no-cpu-baseline --e2e-steps 0 --steps 1 --warmup 0 --streams 1 > gpurun_out/final_ncu_toon.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:reverse_fused -s 1200 -c 1 -o gpurun_out/final_ncu_syn1e7rev -f python bench.py --workload synthetic1e7_rev --no-cpu-baseline --e2e-steps 0 --steps 1 --warmup 0 > gpurun_out/final_ncu_syn.log 2>&1
timeout 900 python bench.py --workload synthetic1e8_rev --steps 2 --e2e-steps 1 > gpurun_out/final_bench_synthetic1e8_rev.json 2> gpurun_out/final_bench_synthetic1e8_rev.err
for f in gpurun_out/final_bench_*.json; do python - $f <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d.get('roofline') or {}; e=d.get('e2e') or {}; c=d.get('cpu_baseline') or {}
    print(sys.argv[1].split('final_bench_')[1], 'value %.3e' % d['value'], 'ms', round(d['ms_per_step'],1), 'e2e %.3e' % (e.get('value') or 0), 'frac', round(r.get('frac',0) or 0,2), 'cpu %.3e' % (c.get('value') or 0), 'parity', d.get('parity_bit_exact_vs_oracle'))
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex)
PY
done
