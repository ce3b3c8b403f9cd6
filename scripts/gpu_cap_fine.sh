B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-variants"
timeout 600 $B > gpurun_out/plain_fine.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:corr_lookup_c32_up2 -s 8 -c 1 -f -o gpurun_out/r02b_fine_tokens_up2 $B > gpurun_out/ncu_fine.log 2>&1
echo rc=$?
