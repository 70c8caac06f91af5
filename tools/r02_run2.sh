# Round 2, GPU run 2: the whole GPU test suite, analysis kernel times, gather ceiling, chain timing, reference arm
mkdir -p gpurun_out
T=${1:-r02b}
( time timeout 2400 python -m pytest tests -m gpu -q ) > gpurun_out/${T}_pytest.log 2>&1
tail -15 gpurun_out/${T}_pytest.log
for a in "20 8192 4 512" "20 8192 4 1024" "20 8192 4 256" "20 8192 8 512" "20 4096 4 512"; do tools/gather_ceiling $a; done > gpurun_out/${T}_gather_ceiling.jsonl 2>&1
cat gpurun_out/${T}_gather_ceiling.jsonl
python tools/analysis_probe.py synthetic1e7_fwd 2 > gpurun_out/${T}_probe_syn1e7.out 2>&1; tail -2 gpurun_out/${T}_probe_syn1e7.out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_analysis_launches_syn1e7.csv python tools/analysis_probe.py synthetic1e7_fwd 1 > gpurun_out/${T}_probe_ncu.out 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(l for l in open('gpurun_out/'+__import__('os').environ.get('T','r02b')+'_analysis_launches_syn1e7.csv') if l.startswith('"'))]
h=rows[0]; ik=h.index("Kernel Name"); iv=h.index("Metric Value")
tot=collections.Counter(); cnt=collections.Counter()
for r in rows[1:]:
    k=r[ik].split('(')[0][-60:]; tot[k]+=float(r[iv].replace(',',''))/1e6; cnt[k]+=1
for k,v in tot.most_common(25): print(f"{v:10.3f} ms {cnt[k]:6d}  {k}")
PY
python tools/analysis_probe.py rosenbrock1e7 2 > gpurun_out/${T}_probe_rosen.out 2>&1; tail -2 gpurun_out/${T}_probe_rosen.out
for spec in "100 100" "20 500"; do adept-2_b200/dropin/_build/chain_check $spec /tmp/chain.bin gpu; adept-2_b200/dropin/_build/chain_check $spec /tmp/chain.bin gpu; done > gpurun_out/${T}_chain_check.out 2>&1
cat gpurun_out/${T}_chain_check.out
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/${T}_bench_cfg4_ref.json 2> gpurun_out/${T}_bench_cfg4_ref.err; cut -c1-400 gpurun_out/${T}_bench_cfg4_ref.json
