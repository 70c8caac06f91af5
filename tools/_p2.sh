mkdir -p gpurun_out
T=r02q
( time timeout 900 python -m pytest tests/test_gpu_resident.py tests/test_gpu_parity.py -m gpu -q -x -k "resident or golden_jacobians or tiny_steps or longer_than or overwritten or non_finite" ) > gpurun_out/${T}_pytest.log 2>&1
tail -6 gpurun_out/${T}_pytest.log
B="--workload toon4096 --no-cpu-baseline --e2e-steps 0 --steps 10 --warmup 3"
for v in "def:" "res:--opt fused_persistent=1" "res_cbs16:--opt fused_persistent=1 --cbs 16" "res_cbs8:--opt fused_persistent=1 --cbs 8" \
         "d512_def:--dirs 512" "d512_res:--dirs 512 --opt fused_persistent=1" "d1024_def:--dirs 1024" "d1024_res:--dirs 1024 --opt fused_persistent=1"; do
  name=${v%%:*}; opts=${v#*:}
  timeout 300 python bench.py $B $opts > gpurun_out/${T}_$name.json 2> gpurun_out/${T}_$name.err
  python - gpurun_out/${T}_$name.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1]); r=d.get('roofline') or {}
    print(sys.argv[1].split('/')[-1], 'ms', round(d['ms_per_step'],2), 'sweep_ms', round(r.get('sweep_ms_per_step',0),2), 'launches', r.get('launches_per_step'), r.get('kernel'), 'parity', d.get('parity_bit_exact_vs_oracle'), d['clocks'].get('sm_mhz'))
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex); print(open(sys.argv[1].replace('.json','.err')).read()[-600:])
PY
done
# in-kernel time stamps (trace build), 512 directions: per-step kernel and resident kernel
export ADEPT_CUDA_LIB=$PWD/adept-2_b200/libadept_cuda_trace.so
ADEPT_CUDA_TRACE_FILE=gpurun_out/${T}_trace_step_512.txt timeout 300 python bench.py $B --dirs 512 --steps 1 --warmup 1 --streams 1 > gpurun_out/${T}_trace_step.json 2> gpurun_out/${T}_trace_step.err
ADEPT_CUDA_TRACE_FILE=gpurun_out/${T}_trace_res_512.txt timeout 300 python bench.py $B --dirs 512 --steps 1 --warmup 1 --opt fused_persistent=1 > gpurun_out/${T}_trace_res.json 2> gpurun_out/${T}_trace_res.err
ADEPT_CUDA_TRACE_FILE=gpurun_out/${T}_trace_res_4096.txt timeout 300 python bench.py $B --steps 1 --warmup 1 --opt fused_persistent=1 > gpurun_out/${T}_trace_res4096.json 2> gpurun_out/${T}_trace_res4096.err
for f in gpurun_out/${T}_trace_*.txt; do echo $f; python tools/trace_summary.py $f; done
