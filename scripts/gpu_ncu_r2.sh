set -x
cd $GRAFT_REPO_ROOT
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 1 --no-secondary --no-cpu-baseline > gpurun_out/launches_r2_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sweep_fw -s 1 -c 1 -o gpurun_out/sweep_fw_r2 python scripts/gpu_ncu_target.py fused 592 > gpurun_out/ncu_sweep_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:perm_philox2 -s 1 -c 1 -o gpurun_out/philox2_r2b python scripts/gpu_ncu_target.py philox 1184 > gpurun_out/ncu_philox2b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:perm_warp -s 1 -c 1 -o gpurun_out/perm_warp_mt_r2 python scripts/gpu_ncu_target.py mt19937 2368 > gpurun_out/ncu_warp_mt.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 400 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "rows_match_reference_golden and (grid8 or grid32 or kat3x3)" > gpurun_out/racecheck_r2.log 2>&1; tail -5 gpurun_out/racecheck_r2.log
