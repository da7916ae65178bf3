#!/bin/bash
# Regenerate the tracked round-2 profile summaries from the artefacts a gpurun evidence pass left in gpurun_out/:
#   bench_final.json                 python bench.py --steps 5 --warmup 3                       (no profiler)
#   launches_final.csv               ncu --metrics gpu__time_duration.sum on a short bench run  (launch list)
#   prof_split_<tag>.ncu-rep         ncu --set full, python scripts/prof_split.py 592 cfg5      (k_scan + k_resolve)
#   prof_split_cfg4_<tag>.ncu-rep    ncu --set full, python scripts/prof_split.py 9472 cfg4
# The captures hold every k_scan / k_resolve launch of prof_split.py (observed-window set-up, warm-up call, profiled call):
# NCU_SKIP=4 keeps the last pair.
# usage: NCU_SKIP=4 scripts/refresh_profiles.sh <tag> "<build note>"
set -e
cd "$(dirname "$0")/.."
TAG=${1:?tag}; NOTE=${2:-}
D=waveform_ot_b200/csrc/wfot_device.cuh; F=waveform_ot_b200/csrc/wfot_fused.cuh; S=waveform_ot_b200/csrc/wfot_split.cu
ln() { grep -n -e "$2" "$1" < /dev/null | head -1 | cut -d: -f1; }
pk=$(ln $D "uint64_t pack2("); fm=$(ln $D "uint64_t fmul2("); tau=$(ln $D "float sqrt_approx("); e64=$(ln $D "void eval64("); hit=$(ln $D "^struct PixelHit"); tm=$(ln $D "void tile_dists(")
ec=$(ln $D "void eval_candidates("); rf=$(ln $D "bool resolve_pixel_flagged("); rfull=$(ln $D "void resolve_pixel_full(")
foot=$(ln $D "^// -* warp footprints"); hot=$(ln $D "^// -* the hot loop"); prep=$(ln $D "window preparation"); epi=$(ln $D "per-pixel epilogue values")
sp=$(ln $F "double store_pixel("); p2=$(ln $F "---------------- P2"); p3=$(ln $F "---------------- P3"); p4=$(ln $F "---------------- P4"); pe=$(ln $F "resident CTAs of a kernel")
ks=$(ln $S "k_scan(FusedArgs a)"); kr=$(ln $S "k_resolve(FusedArgs a)"); hs=$(ln $S "^// -* host side")
GR="tau32:wfot_device.cuh:$tau-$((e64-4)),eval64:wfot_device.cuh:$e64-$((hit-1)),tile_mask (packed FP32 re-evaluation):wfot_device.cuh:$((tm-1))-$((ec-6)),eval_candidates:wfot_device.cuh:$((ec-1))-$((rf-3)),resolve_pixel_flagged:wfot_device.cuh:$((rf-1))-$((rfull-1)),resolve_warp/full:wfot_device.cuh:$rfull-$((foot-1)),prep_window:wfot_device.cuh:$prep-$((epi-1)),store_pixel (density epilogue + slab):wfot_fused.cuh:$((sp-1))-$((sp+36)),P2 column/row sums:wfot_fused.cuh:$p2-$((p3-1)),P3 OT driver:wfot_fused.cuh:$p3-$((p4-1)),P4 gradient assembly:wfot_fused.cuh:$p4-$((pe-1)),warp_ot1d + canon sums:wfot_ot.cuh:1-400,k_resolve body:wfot_split.cu:$kr-$((hs-1))"
GS="tau32:wfot_device.cuh:$tau-$((e64-4)),footmap/lane_block:wfot_device.cuh:$foot-$((hot-1)),scan_block:wfot_device.cuh:$hot-$((prep-1)),prep_window:wfot_device.cuh:$prep-$((epi-1)),k_scan body:wfot_split.cu:$ks-$((kr-12))"
if [ -f gpurun_out/bench_final.json ]; then cp gpurun_out/bench_final.json profiles/r02_bench.json; fi
if [ -f gpurun_out/launches_final.csv ]; then
  cp gpurun_out/launches_final.csv profiles/r02_bench_launches.csv
  python scripts/ncu_launch_summary.py gpurun_out/launches_final.csv 'launch list of `python bench.py --steps 2 --warmup 3 --cpu-sample 0 --secondary 0 --parity-windows 0 --sweep-windows 0` under ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares)' > profiles/r02_bench_launches_summary.txt
fi
for cfg in cfg5 cfg4; do
  if [ $cfg = cfg5 ]; then rep=gpurun_out/prof_split_$TAG.ncu-rep; out=profiles/r02_split_ncu_summary.txt; what="cfg5 shape (nt=1024, 256x256 grid), 592 windows"; cmd="python scripts/prof_split.py 592 cfg5"
  else rep=gpurun_out/prof_split_cfg4_$TAG.ncu-rep; out=profiles/r02_split_cfg4_ncu_summary.txt; what="cfg4 shape (nt=61, 79x61 grid), 9472 windows"; cmd="python scripts/prof_split.py 9472 cfg4"; fi
  [ -f $rep ] || continue
  ncu -i $rep ${NCU_SKIP:+--launch-skip $NCU_SKIP} --page source --csv --print-source cuda,sass > gpurun_out/src_split_$cfg.csv 2>/dev/null
  { echo "ncu --set full --clock-control none --import-source on, k_scan + k_resolve (two-kernel form of the fused path), $what; $NOTE"
    echo "command: $cmd   (report: $rep, not tracked)"; echo
    python scripts/ncu_key.py $rep 2>/dev/null; echo
    echo "k_resolve: warp-state samples / executed instructions by source region (scripts/ncu_src.py):"
    NCU_GROUPS="$GR" python scripts/ncu_src.py gpurun_out/src_split_$cfg.csv k_resolve 12; echo
    echo "k_scan: warp-state samples / executed instructions by source region:"
    NCU_GROUPS="$GS" python scripts/ncu_src.py gpurun_out/src_split_$cfg.csv k_scan 12; } > $out
done
python - "$TAG" <<'PY'
import csv, io, json, subprocess, sys
rep = "gpurun_out/prof_split_%s.ncu-rep" % sys.argv[1]
import os
skip = ["--launch-skip", os.environ["NCU_SKIP"]] if os.environ.get("NCU_SKIP") else []
out = subprocess.run(["ncu", "-i", rep] + skip + ["--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
per = {}
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    k = "k_scan" if "k_scan" in name else "k_resolve" if "k_resolve" in name else None
    if not k:
        continue
    rd = float(r[hdr.index("dram__bytes_read.sum")]) * mul[units[hdr.index("dram__bytes_read.sum")]]
    wr = float(r[hdr.index("dram__bytes_write.sum")]) * mul[units[hdr.index("dram__bytes_write.sum")]]
    per[k] = {"read": rd / 592, "write": wr / 592, "kernel": name}
d = {"kernel": "k_scan + k_resolve (two-kernel form), one ncu --set full capture of 592 cfg5 windows each", "round": 2,
     "dram_bytes_per_window": sum(v["read"] + v["write"] for v in per.values()), "per_kernel": per,
     "source": "%s (dram__bytes_read.sum + dram__bytes_write.sum), summarised in profiles/r02_split_ncu_summary.txt" % rep,
     "windows_in_capture": 592}
json.dump(d, open("profiles/traffic.json", "w"), indent=1)
print("traffic per window", d["dram_bytes_per_window"])
PY
