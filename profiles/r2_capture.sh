#!/bin/bash
# Round-2 profile artifacts (run on the GPU box through gpurun): launch list of the C2 step + ncu --set full of the
# kernels the roofline block names.  Outputs under gpurun_out/; summaries are produced here with profiles/ncu_summary.py.
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-retrieval --no-extras"
$B > gpurun_out/r2_cap_plain.json 2> gpurun_out/r2_cap_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_c2.csv $B > /dev/null 2>&1
for k in spmm_fwd_v4 gemm_tc3_tma_kernel dw_gather_v4 bn_stats_kernel gemm_tc3_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o gpurun_out/r2_full_$k $B > gpurun_out/r2_full_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
