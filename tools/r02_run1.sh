# Round 2, GPU run 1 (under gpurun, one GPU): tests, the default bench (config 4 full size), config 2, the reference arm,
# and whole-sweep / stratified DRAM-traffic captures.
mkdir -p gpurun_out
T=${1:-r02a}
( time timeout 1800 python -m pytest tests -m gpu -q -x ) > gpurun_out/${T}_pytest.log 2>&1
tail -5 gpurun_out/${T}_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench_cfg4.json 2> gpurun_out/${T}_bench_cfg4.err
tail -c 600 gpurun_out/${T}_bench_cfg4.err
timeout 400 python bench.py --workload toon4096 > gpurun_out/${T}_bench_toon4096.json 2> gpurun_out/${T}_bench_toon4096.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/${T}_bench_cfg4_ref.json 2> gpurun_out/${T}_bench_cfg4_ref.err
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
for spec in "synthetic1e7_fwd 0 1" "synthetic1e7_fwd 1 1" "synthetic1e8_fwd 0 50" "synthetic1e8_fwd 1 50"; do
  set -- $spec
  tag=${1}_m${2}
  rm -f gpurun_out/${T}_prof_$tag.log
  ADEPT_CUDA_PROFILE_LOG=gpurun_out/${T}_prof_$tag.log timeout 900 ncu --profile-from-start off --clock-control none --metrics $M --csv \
     --log-file gpurun_out/${T}_traffic_$tag.csv python tools/ncu_traffic.py run $1 $2 $3 > gpurun_out/${T}_traffic_$tag.out 2>&1
  tail -2 gpurun_out/${T}_traffic_$tag.out
  python tools/ncu_traffic.py summarise gpurun_out/${T}_traffic_$tag.csv gpurun_out/${T}_prof_$tag.log $2 $tag > gpurun_out/${T}_traffic_$tag.json 2>> gpurun_out/${T}_traffic_$tag.out
  cat gpurun_out/${T}_traffic_$tag.json
done
for f in gpurun_out/${T}_bench_*.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1]); r=d.get('roofline') or {}; e=d.get('e2e') or {}; c=d.get('cpu_baseline') or {}
    o=d.get('reverse') or d.get('forward') or {}
    print(sys.argv[1].split('_bench_')[1], 'value %.3e' % d['value'], 'ms', round(d['ms_per_step'],1), 'e2e %.3e' % (e.get('value') or 0), 'frac', round(r.get('frac',0) or 0,3),
          'cpu %.3e' % (c.get('value') or 0), 'parity', d.get('parity_bit_exact_vs_oracle'), 'other', o.get('ms_per_step'), (o.get('roofline') or {}).get('frac'))
    print('   phases', (e.get('phases_ms') or {}))
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex)
PY
done
