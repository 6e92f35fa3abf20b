#!/bin/bash
# Runs on the GPU box (gpurun): tests, both bench arms, the ncu launch list of the bench command and --set full captures
# of the hot kernels.  Everything lands in gpurun_out/<tag>_*; tools/round_profiles_post.sh turns it into profiles/.
tag=${1:-r2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${tag}_launches_bench.csv \
  python bench.py --steps 2 --warmup 1 --skip-e2e --skip-cpu --skip-pose-opt > gpurun_out/${tag}_bench_under_ncu.log 2>&1; echo "launchlist rc=$?"
for k in k_update_tma:2:3 k_accumulate_runs:1:1 k_potrf_smem:20:1 k_trsm_ll:20:1; do
  IFS=: read name skip cnt <<< "$k"
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$name --launch-skip $skip --launch-count $cnt -f -o gpurun_out/${tag}_full_$name \
    python tools/prof_iteration.py --iters 1 > gpurun_out/${tag}_full_$name.log 2>&1; echo "full $name rc=$?"
done
timeout 300 python tools/level_profile.py > gpurun_out/${tag}_levels.txt 2>&1
timeout 300 python tools/upd_experiment.py > gpurun_out/${tag}_experiments.txt 2>&1
bash tools/round_profiles_post.sh $tag gpurun_out/profiles > gpurun_out/${tag}_post.log 2>&1; echo "post rc=$?"
rm -f gpurun_out/*.ncu-rep          # summarised above; too big for the 64 MiB that travel back
du -sh gpurun_out
tail -c 600 gpurun_out/${tag}_bench_1gpu.json
