#!/bin/bash
# Round-2 profiling pass (run on the GPU box through gpurun; outputs under gpurun_out/):
#   1. ncu launch list of the timed steps of the bench command (NVTX range "fino_timed")
#   2. ncu --set full of one launch each: the d=128 attention kernel (DRAM traffic), the VAE conv, the VAE RMS-norm
#      kernel, the CogVideoX LayerNorm(64)+RoPE kernel
#   3. ncu --set full of the five Wan GEMM shapes, this library and cuBLAS back to back in one session
set -x
O=gpurun_out
NCU="ncu --clock-control none"
$NCU --nvtx --nvtx-include "fino_timed/" --metrics gpu__time_duration.sum --csv --log-file $O/r02_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > $O/r02_launches_bench.log 2>&1
$NCU --set full --import-source on -k regex:attn_fwd_kernel --launch-skip 2 -c 1 -o $O/r02_attn_d128 -f python tools/prof_kernels.py attn > /dev/null 2>&1
$NCU --set full --import-source on -k regex:conv_cl_kernel --launch-skip 2 -c 1 -o $O/r02_vae_conv -f python tools/prof_kernels.py conv > /dev/null 2>&1
$NCU --set full -k regex:rms_act_cl_kernel --launch-skip 2 -c 1 -o $O/r02_vae_rms_act -f python tools/prof_kernels.py conv > /dev/null 2>&1
$NCU --set full -k regex:qk_ln64_rope_kernel --launch-skip 2 -c 1 -o $O/r02_qk_ln64 -f python tools/prof_kernels.py ln64 > /dev/null 2>&1
$NCU --set full --profile-from-start off -k regex:"gemm2_bf16_kernel|nvjet|cutlass|gemm_fixup" -o $O/r02_gemm5 -f python tools/prof_kernels.py gemm5 > $O/r02_gemm5.log 2>&1
ls -la $O/*.ncu-rep
