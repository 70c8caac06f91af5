# Round 2, final single-GPU measurement run (under gpurun): tests, every bench line archived under profiles/, traffic
# captures of the default workload, one ncu --set full capture of the dominant kernel, launch list.
mkdir -p gpurun_out
T=${1:-r02z}
( time timeout 2400 python -m pytest tests -m gpu -q ) > gpurun_out/${T}_pytest.log 2>&1
tail -4 gpurun_out/${T}_pytest.log
# the same suite against the build with our own device-side bounds checks (make -C adept-2_b200/csrc CHECK=1)
( time ADEPT_CUDA_LIB=$PWD/adept-2_b200/libadept_cuda_check.so timeout 2400 python -m pytest tests -m gpu -q -x ) > gpurun_out/${T}_pytest_check.log 2>&1
tail -2 gpurun_out/${T}_pytest_check.log
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench_synthetic1e8_fwd.json 2> gpurun_out/${T}_bench_synthetic1e8_fwd.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/${T}_bench_synthetic1e8_fwd_ref.json 2> gpurun_out/${T}_bench_synthetic1e8_fwd_ref.err
timeout 600 python bench.py --workload toon4096 > gpurun_out/${T}_bench_toon4096.json 2> gpurun_out/${T}_bench_toon4096.err
timeout 600 python bench.py --workload lw100 > gpurun_out/${T}_bench_lw100.json 2> gpurun_out/${T}_bench_lw100.err
timeout 600 python bench.py --workload synthetic1e7_fwd > gpurun_out/${T}_bench_synthetic1e7_fwd.json 2> gpurun_out/${T}_bench_synthetic1e7_fwd.err
timeout 900 python bench.py --workload rosenbrock1e7 --steps 5 > gpurun_out/${T}_bench_rosenbrock1e7.json 2> gpurun_out/${T}_bench_rosenbrock1e7.err
timeout 900 python bench.py --workload ensemble1e5 --steps 3 > gpurun_out/${T}_bench_ensemble1e5.json 2> gpurun_out/${T}_bench_ensemble1e5.err
timeout 900 python bench.py --workload checkpoint --steps 2 > gpurun_out/${T}_bench_checkpoint.json 2> gpurun_out/${T}_bench_checkpoint.err
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
for spec in "synthetic1e8_fwd 0 50" "synthetic1e8_fwd 1 50"; do
  set -- $spec
  tag=${1}_m${2}
  rm -f gpurun_out/${T}_prof_$tag.log
  ADEPT_CUDA_PROFILE_LOG=gpurun_out/${T}_prof_$tag.log timeout 900 ncu --profile-from-start off --clock-control none --metrics $M --csv \
     --log-file gpurun_out/${T}_traffic_$tag.csv python tools/ncu_traffic.py run $1 $2 $3 > gpurun_out/${T}_traffic_$tag.out 2>&1
  python tools/ncu_traffic.py summarise gpurun_out/${T}_traffic_$tag.csv gpurun_out/${T}_prof_$tag.log $2 $tag > gpurun_out/${T}_traffic_$tag.json 2>> gpurun_out/${T}_traffic_$tag.out
  cat gpurun_out/${T}_traffic_$tag.json
done
# one full capture of the dominant kernel (a wide step in the middle of the forward sweep of the 1e7 tape: same shape, 1/10 length)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:forward_balanced -s 1254 -c 1 -o gpurun_out/${T}_ncu_forward_balanced -f \
   python bench.py --workload synthetic1e7_fwd --no-cpu-baseline --e2e-steps 0 --reverse-steps 0 --steps 1 --warmup 3 --opt wide_streams=1 > gpurun_out/${T}_ncu_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_launches_synthetic1e7_fwd.csv \
   python bench.py --workload synthetic1e7_fwd --no-cpu-baseline --e2e-steps 0 --reverse-steps 0 --steps 1 --warmup 3 > gpurun_out/${T}_ncu_launches.log 2>&1
for f in gpurun_out/${T}_bench_*.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1]); r=d.get('roofline') or {}; e=d.get('e2e') or {}; c=d.get('cpu_baseline') or {}
    o=d.get('reverse') or d.get('forward') or {}
    print(sys.argv[1].split('_bench_')[1], 'value %.3e' % (d.get('value') or 0), 'ms', round(d.get('ms_per_step') or 0,2), 'e2e %.3e' % (e.get('value') or 0), 'e2e_ms', e.get('ms_per_step'), 'frac', round(r.get('frac',0) or 0,3),
          'cpu %.3e' % (c.get('value') or 0), 'parity', d.get('parity_bit_exact_vs_oracle', d.get('parity_bit_exact_vs_reference_every_member')), 'other', o.get('ms_per_step'), (o.get('roofline') or {}).get('frac'))
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex)
PY
done
