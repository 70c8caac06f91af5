mkdir -p gpurun_out
run() { name=$1; shift; timeout 300 python bench.py --no-cpu-baseline "$@" > gpurun_out/t3_$name.json 2> gpurun_out/t3_$name.err; python - gpurun_out/t3_$name.json $name <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d.get('roofline') or {}
    e=d.get('e2e') or {}
    print(sys.argv[2], 'ms', round(d['ms_per_step'],1), 'parity', d.get('parity_bit_exact_vs_oracle'), 'launch_us', round(r.get('avg_launch_us',0),1), 'quarters', [round(x,1) for x in (r.get('ms_by_quarter_of_sweep') or []) if x is not None], 'e2e', e.get('value'), {k:(round(v,1) if isinstance(v,float) else v) for k,v in (e.get('phases') or {}).items()})
except Exception as ex:
    print(sys.argv[2], 'FAILED', ex)
PY
}
run toon_auto --e2e-steps 2
run toon_p2048 --pass-dirs 2048 --e2e-steps 0
run toon_p1024 --pass-dirs 1024 --e2e-steps 0
run toon_p1024_c8 --pass-dirs 1024 --cbs 8 --chunk 32 --e2e-steps 0
run toon_w16k --window 16384 --e2e-steps 2
run toon_w8k --window 8192 --e2e-steps 2
for skip in 300 1500 2900; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:reverse_fused -s $skip -c 1 -o gpurun_out/t3_ncu_toon_s$skip -f python bench.py --no-cpu-baseline --e2e-steps 0 --steps 1 --warmup 0 --streams 1 > gpurun_out/t3_ncu_$skip.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
