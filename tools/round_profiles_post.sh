#!/bin/bash
# Turns the gpurun_out/<tag>_* artefacts of tools/round_profiles.sh into the tracked summaries under profiles/.
tag=${1:-r2}
out=${2:-profiles}          # on the GPU box: gpurun_out/profiles (only gpurun_out/ travels back; the .ncu-rep files are too big to travel)
cd "$(dirname "$0")/.."
mkdir -p $out
{
  echo "# ncu launch list of: ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv python bench.py --steps 2 --warmup 1 --skip-e2e --skip-cpu --skip-pose-opt"
  echo "# config 2 (300 frames, 384x224, grid 16x12, matchSeparation 10). Cold-cache, serialised per-launch times: compare SHARES with bench.py's"
  echo "# linear_kernels_ms_serialised / roofline.share_of_step_serialised, not absolutes. k_dmma_peak is the fp64 tensor peak probe (not part of a step)."
  python tools/launch_summary.py gpurun_out/${tag}_launches_bench.csv 30
} > $out/${tag}_launches_bench_config2.txt
for k in k_update_tma k_accumulate_runs k_potrf_smem k_trsm_ll; do
  {
    echo "# ncu --set full --clock-control none --import-source on -k regex:$k (python tools/prof_iteration.py --iters 1, config 2)"
    python tools/ncu_metrics.py gpurun_out/${tag}_full_$k.ncu-rep
    echo; echo "## top SASS instructions by warp-stall samples"
    python tools/ncu_hot.py gpurun_out/${tag}_full_$k.ncu-rep $k 16
  } > $out/${tag}_ncu_$k.txt 2>&1
done
python - "$tag" "$out" <<'PY'
import csv, json, subprocess, sys
tag = sys.argv[1]
out = {"source": f"ncu --set full --clock-control none captures of tools/prof_iteration.py --iters 1 (config 2), tools/round_profiles.sh, gpurun_out/{tag}_full_*.ncu-rep; mean over the captured launches"}
for k in ("k_update_tma", "k_accumulate_runs"):
    txt = subprocess.run(["ncu", "-i", f"gpurun_out/{tag}_full_{k}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines())); H = rows[0]
    rd, wr, du = H.index("dram__bytes_read.sum"), H.index("dram__bytes_write.sum"), H.index("gpu__time_duration.sum")
    units = rows[1]
    def b(v, u): return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    vals = [b(r[rd], units[rd]) + b(r[wr], units[wr]) for r in rows[2:]]
    out[k] = {"dram_bytes_per_launch": sum(vals) / len(vals), "launches_captured": len(vals), "per_launch": vals,
              "duration_us": [float(r[du]) * {"us": 1, "ms": 1e3, "ns": 1e-3}.get(units[du], 1) for r in rows[2:]]}
json.dump(out, open(f"{sys.argv[2]}/{tag}_ncu_traffic.json", "w"), indent=1)
PY
{
  echo "# cuobjdump -sass robust_cvd_b200/librcvd_b200.so, function k_update_tma<1>: the instructions that prove the Blackwell data path"
  echo "# (UTMALDG = cp.async.bulk.tensor, SYNCS = mbarrier, DMMA = fp64 tensor core; fp64 has no tcgen05 kind)."
  cuobjdump -sass robust_cvd_b200/librcvd_b200.so 2>/dev/null | awk '/Function : .*k_update_tmaILi1/{f=1} f{print} /Function : /{if(f&&!/k_update_tmaILi1/)exit}' > /tmp/sass_upd.txt
  echo "# counts: $(grep -c UTMALDG /tmp/sass_upd.txt) UTMALDG, $(grep -c 'SYNCS' /tmp/sass_upd.txt) SYNCS (mbarrier), $(grep -c 'DMMA' /tmp/sass_upd.txt) DMMA.8x8x4, $(grep -c 'LDS.64' /tmp/sass_upd.txt) LDS.64, $(grep -c 'LDGSTS' /tmp/sass_upd.txt) LDGSTS, $(grep -c 'LD.E.64' /tmp/sass_upd.txt) generic LD.E.64"
  grep -n "UTMALDG\|SYNCS\|UBLKCP" /tmp/sass_upd.txt | sed 's/\/\* 0x[0-9a-f]* \*\///' | head -40
  echo "# whole library:"
  cuobjdump -sass robust_cvd_b200/librcvd_b200.so 2>/dev/null > /tmp/sass_all.txt
  for m in UTMALDG SYNCS DMMA LDGSTS 'RED.E.ADD.F64' 'REDG.E.ADD.F64' UTCHMMA LDTM; do echo "#   $m: $(grep -c "$m" /tmp/sass_all.txt)"; done
} > $out/sass_k_update_tma.txt
cp gpurun_out/${tag}_bench_1gpu.json $out/${tag}_bench_1gpu.json
cp gpurun_out/${tag}_bench_reference.json $out/${tag}_bench_reference_arm.json
cp gpurun_out/${tag}_levels.txt $out/${tag}_factor_levels_config2.txt
cp gpurun_out/${tag}_experiments.txt $out/${tag}_kernel_variants_ab.txt
tail -3 gpurun_out/${tag}_pytest_gpu.log > $out/${tag}_pytest_gpu.txt
