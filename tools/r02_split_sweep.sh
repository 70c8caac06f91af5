# Config 2 (Toon nx=4096 nt=1000, reverse) with few directions per device: chunks shared by 1/2/4/8 CTAs ("split"), one launch per
# step and the resident kernel ("fused_persistent"), plus in-kernel time stamps from the trace build (make -C adept-2_b200/csrc TRACE=1).
# Produced profiles/r02_split_*.json and profiles/r02_fused_trace_summary.txt.  NOTE: bound the pytest line (it once ran into a hang).
mkdir -p gpurun_out
T=r02s
( time timeout 120 python -m pytest tests/test_gpu_resident.py tests/test_gpu_parity.py -m gpu -q -x -k "resident or golden_jacobians or tiny_steps or longer_than or overwritten or non_finite" ) > gpurun_out/${T}_pytest.log 2>&1
tail -6 gpurun_out/${T}_pytest.log
B="--workload toon4096 --no-cpu-baseline --e2e-steps 0 --steps 8 --warmup 3"
run() { name=$1; shift
  timeout 300 python bench.py $B "$@" > gpurun_out/${T}_$name.json 2> gpurun_out/${T}_$name.err
  python - gpurun_out/${T}_$name.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1]); r=d.get('roofline') or {}
    print(sys.argv[1].split('/')[-1], 'ms', round(d['ms_per_step'],2), 'sweep_ms', round(r.get('sweep_ms_per_step',0),2), 'launches', r.get('launches_per_step'), (r.get('kernel') or '')[:24], 'parity', d.get('parity_bit_exact_vs_oracle'), d['clocks'].get('sm_mhz'))
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex); print(open(sys.argv[1].replace('.json','.err')).read()[-600:])
PY
}
for D in 512 1024 2048; do
  for sp in 1 2 4 8; do
    run d${D}_step_s$sp --dirs $D --opt split=$sp
    run d${D}_res_s$sp --dirs $D --opt split=$sp --opt fused_persistent=1
  done
done
run d4096_step_s1 --opt split=1
run d4096_step_s2 --opt split=2
run d4096_res_s1_cbs8 --opt split=1 --opt fused_persistent=1 --cbs 8
run d4096_res_s2_cbs8 --opt split=2 --opt fused_persistent=1 --cbs 8
run d512_step_s8_sub16 --dirs 512 --opt split=8 --opt sub=16
run d512_res_s16_sub4 --dirs 512 --opt split=16 --opt sub=4 --opt fused_persistent=1
export ADEPT_CUDA_LIB=$PWD/adept-2_b200/libadept_cuda_trace.so
ADEPT_CUDA_TRACE_FILE=gpurun_out/${T}_trace_step_512.txt timeout 300 python bench.py $B --dirs 512 --steps 1 --warmup 1 --streams 1 --opt split=8 > gpurun_out/${T}_trace_step.json 2> gpurun_out/${T}_trace_step.err
ADEPT_CUDA_TRACE_FILE=gpurun_out/${T}_trace_res_512.txt timeout 300 python bench.py $B --dirs 512 --steps 1 --warmup 1 --opt fused_persistent=1 --opt split=8 > gpurun_out/${T}_trace_res.json 2> gpurun_out/${T}_trace_res.err
for f in gpurun_out/${T}_trace_*.txt; do echo $f; python tools/trace_summary.py $f; done
