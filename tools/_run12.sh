mkdir -p gpurun_out
for i in 1 2; do
timeout 300 python bench.py --no-cpu-baseline --e2e-steps 6 --steps 3 > gpurun_out/t12_$i.json 2> gpurun_out/t12_$i.err
python - gpurun_out/t12_$i.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); e=d['e2e']
print('ms', round(d['ms_per_step'],1), 'e2e_ms', round(e['ms_per_step'],1), e['phases_ms']['each_step_upload_jacobian_replay_ms'])
PY
done
nproc; uptime
