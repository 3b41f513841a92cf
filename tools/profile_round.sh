#!/bin/bash
# ncu evidence for a round (run under gpurun; summaries are written by tools/ncu_summary.py on the CPU box).
set -x
TAG=${1:-r01c}
# (1) launch list of the timed steps: bench.py brackets its timed region with cudaProfilerStart/Stop; kernels are launched
#     from the host (--no-graphs) so that every launch is listed by name
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-generation --no-mar --no-extras --no-config5 --no-gpu-reference --no-graphs > gpurun_out/ncu_launches_${TAG}.log 2>&1
tail -1 gpurun_out/ncu_launches_${TAG}.log | cut -c1-200
# (1b) the same for one HMA-MAR training step + 2 ancestral sampler steps
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_mar_launches.csv \
    python tools/mar_profile.py --sample-steps 2 > gpurun_out/ncu_mar_launches_${TAG}.log 2>&1
tail -1 gpurun_out/ncu_mar_launches_${TAG}.log | cut -c1-200
if [ "$2" != "launches-only" ]; then
# (2) full capture of the hot kernels at the benchmark shapes (2 layers: same shapes per launch)
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:'gemm_nt_kernel|attn_spatial|attn_temporal_tc|gemm_wgrad|ln_bwd|ln_fwd' -c 46 -o gpurun_out/${TAG}_prof_hot \
    python bench.py --layers 2 --steps 1 --warmup 3 --no-cpu-baseline --no-generation --no-mar --no-extras --no-config5 --no-gpu-reference --no-graphs > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -1 gpurun_out/ncu_full_${TAG}.log | cut -c1-200
# (2b) full capture of the STMAR row-wise kernels and of the diffusion-MLP GEMM shapes
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:'mar_|dropout' -c 40 -o gpurun_out/${TAG}_prof_mar \
    python tools/mar_profile.py --layers 1 --sample-steps 1 > gpurun_out/ncu_full_mar_${TAG}.log 2>&1
tail -1 gpurun_out/ncu_full_mar_${TAG}.log | cut -c1-200
fi
