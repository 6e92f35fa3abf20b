#!/bin/bash
# Reduced end-of-round capture (GPU-minute budget): tests, bench, per-level profile, ncu launch list and --set full captures of the
# kernels changed last (k_potrf_smem, k_trsm_ll, k_substitution).  Summaries land in gpurun_out/profiles (the .ncu-rep files do not travel).
tag=${1:-r2b}
mkdir -p gpurun_out/profiles
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/profiles/${tag}_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/profiles/${tag}_pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.err; echo "bench rc=$?"
grep '^{' gpurun_out/${tag}_bench_1gpu.json | tail -1 > gpurun_out/profiles/${tag}_bench_1gpu.json
timeout 200 python tools/level_profile.py > gpurun_out/profiles/${tag}_factor_levels_config2.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${tag}_launches_bench.csv \
  python bench.py --steps 2 --warmup 1 --skip-e2e --skip-cpu --skip-pose-opt > gpurun_out/${tag}_bench_under_ncu.log 2>&1; echo "launchlist rc=$?"
{
  echo "# ncu launch list of: ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv python bench.py --steps 2 --warmup 1 --skip-e2e --skip-cpu --skip-pose-opt"
  echo "# config 2, end of round 2 (after the pivot-tile / chain-warp / substitution work). Cold-cache, serialised per-launch times: compare SHARES, not absolutes."
  python tools/launch_summary.py gpurun_out/${tag}_launches_bench.csv 30
} > gpurun_out/profiles/${tag}_launches_bench_config2.txt 2>&1
for k in k_potrf_smem:20:1 k_trsm_ll:20:1 k_substitution:0:1; do
  IFS=: read name skip cnt <<< "$k"
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$name --launch-skip $skip --launch-count $cnt -f -o gpurun_out/${tag}_full_$name \
    python tools/prof_iteration.py --iters 1 > gpurun_out/${tag}_full_$name.log 2>&1; echo "full $name rc=$?"
  {
    echo "# ncu --set full --clock-control none --import-source on -k regex:$name (python tools/prof_iteration.py --iters 1, config 2), end of round 2"
    python tools/ncu_metrics.py gpurun_out/${tag}_full_$name.ncu-rep
    echo; echo "## top SASS instructions by warp-stall samples"
    python tools/ncu_hot.py gpurun_out/${tag}_full_$name.ncu-rep $name 16
  } > gpurun_out/profiles/${tag}_ncu_$name.txt 2>&1
done
rm -f gpurun_out/*.ncu-rep
du -sh gpurun_out
tail -c 1500 gpurun_out/profiles/${tag}_bench_1gpu.json
