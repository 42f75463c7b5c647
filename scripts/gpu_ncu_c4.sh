set -x
cd $GRAFT_REPO_ROOT
export PZ_PIPELINE=0
PZ_CTA_WARPS=4 PZ_G32_CTAS=6 ncu --set full --clock-control none -k regex:sweep_cta -s 1 -c 1 -o gpurun_out/c4_888x4_r2 python scripts/gpu_ncu_target.py fused 888 1024 > gpurun_out/ncu_c4_a.log 2>&1
PZ_CTA_WARPS=32 PZ_G32_CTAS=1 ncu --set full --clock-control none -k regex:sweep_cta -s 1 -c 1 -o gpurun_out/c4_148x32_r2 python scripts/gpu_ncu_target.py fused 148 1024 > gpurun_out/ncu_c4_b.log 2>&1
PZ_CTA_WARPS=32 PZ_G32_CTAS=1 ncu --set full --clock-control none -k regex:sweep_cta -s 1 -c 1 -o gpurun_out/c4_24x32_r2 python scripts/gpu_ncu_target.py fused 24 1024 > gpurun_out/ncu_c4_c.log 2>&1
ls -la gpurun_out/c4_*.ncu-rep
