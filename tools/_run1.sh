set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/t1_pytest.log 2>&1
tail -5 gpurun_out/t1_pytest.log
timeout 300 python bench.py --no-cpu-baseline --fused 1 > gpurun_out/t1_bench_toon_fused.json 2> gpurun_out/t1_bench_toon_fused.err
timeout 300 python bench.py --no-cpu-baseline --fused 0 > gpurun_out/t1_bench_toon_unfused.json 2> gpurun_out/t1_bench_toon_unfused.err
timeout 300 python bench.py --no-cpu-baseline --workload synthetic1e7_rev --fused 1 > gpurun_out/t1_bench_syn_fused.json 2> gpurun_out/t1_bench_syn_fused.err
timeout 300 python bench.py --no-cpu-baseline --workload synthetic1e7_rev --fused 0 > gpurun_out/t1_bench_syn_unfused.json 2> gpurun_out/t1_bench_syn_unfused.err
cat gpurun_out/t1_bench_*.json | cut -c1-600
