mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/t13_pytest.log 2>&1
tail -4 gpurun_out/t13_pytest.log
run() { name=$1; shift; timeout 300 python bench.py --no-cpu-baseline --e2e-steps 0 "$@" > gpurun_out/t13_$name.json 2> gpurun_out/t13_$name.err; python - gpurun_out/t13_$name.json $name <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d.get('roofline') or {}
    print(sys.argv[2], 'ms', round(d['ms_per_step'],1), 'parity', d.get('parity_bit_exact_vs_oracle'), 'launch_us', round(r.get('avg_launch_us',0),1), 'launches', d.get('gpu_launches'), 'frac', round(r.get('frac',0),2))
except Exception as ex:
    print(sys.argv[2], 'FAILED', ex)
PY
}
run toon_tail1
run toon_tail0 --tail 0
run toon_tail1_b
run toon_tail0_b --tail 0
run toon_512_tail1 --dirs 512
run toon_512_tail0 --dirs 512 --tail 0
run syn_auto --workload synthetic1e7_rev
