#!/bin/bash
# Turns the gpurun_out/<tag>_* artefacts of tools/round_profiles.sh into the tracked summaries under profiles/.
tag=${1:-r1}
cd "$(dirname "$0")/.."
{
  echo "# ncu launch list of: ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv python bench.py --steps 2 --warmup 1 --skip-e2e --skip-cpu"
  echo "# config 2 (300 frames, 384x224, grid 16x12, matchSeparation 10). Cold-cache, serialised per-launch times: compare SHARES with bench.py's"
  echo "# linear_kernels_ms_serialised / roofline.share_of_step_serialised, not absolutes. k_dmma_peak is the fp64 tensor peak probe (not part of a step)."
  python tools/launch_summary.py gpurun_out/${tag}_launches_bench.csv 24
} > profiles/${tag}_launches_bench_config2.txt
for k in k_gemm_nt k_accumulate_fast k_potrf_smem k_trsm_ll; do
  {
    echo "# ncu --set full --clock-control none --import-source on -k regex:$k (python tools/prof_iteration.py --iters 1, config 2)"
    python tools/ncu_metrics.py gpurun_out/${tag}_full_$k.ncu-rep
    echo; echo "## top SASS instructions by warp-stall samples"
    python tools/ncu_hot.py gpurun_out/${tag}_full_$k.ncu-rep $k 16
  } > profiles/${tag}_ncu_$k.txt 2>&1
done
cp gpurun_out/${tag}_bench_1gpu.json profiles/${tag}_bench_1gpu.json
cp gpurun_out/${tag}_bench_reference.json profiles/${tag}_bench_reference_arm.json
tail -3 gpurun_out/${tag}_pytest_gpu.log > profiles/${tag}_pytest_gpu.txt
