mkdir -p gpurun_out
for skip in 1500 2802; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:reverse_fused -s $skip -c 3 -o gpurun_out/t4_ncu_toon_s$skip -f python bench.py --no-cpu-baseline --e2e-steps 0 --steps 1 --warmup 0 --streams 1 > gpurun_out/t4_ncu_$skip.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
