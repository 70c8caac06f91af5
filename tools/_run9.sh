mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/t11_pytest.log 2>&1
tail -4 gpurun_out/t11_pytest.log
run() { name=$1; shift; timeout 300 python bench.py --no-cpu-baseline --e2e-steps 3 --steps 2 "$@" > gpurun_out/t11_$name.json 2> gpurun_out/t11_$name.err; python - gpurun_out/t11_$name.json $name <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); e=d.get('e2e') or {}; ph=e.get('phases_ms') or {}
    print(sys.argv[2], 'ms', round(d['ms_per_step'],1), 'parity', d.get('parity_bit_exact_vs_oracle'), 'e2e_ms', round(e.get('ms_per_step',0),1), {k:round(v,1) for k,v in ph.items() if v is not None and k in ('upload_ms','analysis_ms','an_window_ms','an_forward_ms','an_fused_ms','replay_ms','d2h_ms')}, 'levels', d['config'].get('steps_in_schedule', d['config'].get('n_levels')))
except Exception as ex:
    print(sys.argv[2], 'FAILED', ex)
PY
}
run toon_w64k
run toon_w32k --window 32768
run toon_w64k_b
