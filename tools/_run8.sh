mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/t8_pytest.log 2>&1
tail -4 gpurun_out/t8_pytest.log
run() { name=$1; shift; timeout 300 python bench.py --no-cpu-baseline --e2e-steps 0 "$@" > gpurun_out/t8_$name.json 2> gpurun_out/t8_$name.err; python - gpurun_out/t8_$name.json $name <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d.get('roofline') or {}
    print(sys.argv[2], 'ms', round(d['ms_per_step'],1), 'parity', d.get('parity_bit_exact_vs_oracle'), 'launch_us', round(r.get('avg_launch_us',0),1), 'quarters', [round(x,1) for x in (r.get('ms_by_quarter_of_sweep') or []) if x is not None], 'frac', round(r.get('frac',0),2))
except Exception as ex:
    print(sys.argv[2], 'FAILED', ex)
PY
}
run toon_auto
run toon_auto_b
run toon_c32_s2 --cbs 32 --streams 2
run toon_c8_s4 --cbs 8 --streams 4
run toon_c16_s4_k32 --chunk 32
run toon_c16_s4_k48 --chunk 48
run syn_auto --workload synthetic1e7_rev
