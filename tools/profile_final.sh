#!/bin/bash
# Round-end ncu evidence for the final kernels (run under gpurun): launch list of the timed steps + full capture of the backward
# hot kernels (attn_spatial_bwd is the roofline kernel) and of the forward ones.
set -x
TAG=${1:-r02b}
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-generation --no-mar --no-extras --no-config5 --no-gpu-reference --no-graphs > gpurun_out/ncu_launches_${TAG}.log 2>&1
tail -1 gpurun_out/ncu_launches_${TAG}.log | cut -c1-160
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:'attn_spatial_bwd|gemm_wgrad_grouped|ln_bwd|attn_temporal_tc_bwd|gemm_nt_kernel' -c 24 -o gpurun_out/${TAG}_prof_bwd \
    python bench.py --layers 1 --steps 1 --warmup 3 --no-cpu-baseline --no-generation --no-mar --no-extras --no-config5 --no-gpu-reference --no-graphs > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -1 gpurun_out/ncu_full_${TAG}.log | cut -c1-160
ncu -i gpurun_out/${TAG}_prof_bwd.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_bwd_raw.csv
ls -la gpurun_out/ | tail -8
