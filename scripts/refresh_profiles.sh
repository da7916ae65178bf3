#!/bin/bash
# Regenerate the tracked profile summaries from the artefacts a gpurun evidence pass left in gpurun_out/:
#   bench_final.json, launches_final.csv, prof_fused_<tag>.ncu-rep, prof_fused_cfg4_<tag>.ncu-rep
# usage: scripts/refresh_profiles.sh <tag> "<build note>"
set -e
cd "$(dirname "$0")/.."
TAG=${1:?tag}; NOTE=${2:-}
D=waveform_ot_b200/csrc/wfot_device.cuh; F=waveform_ot_b200/csrc/wfot_fused.cu
ln() { grep -n -e "$2" "$1" < /dev/null | head -1 | cut -d: -f1; }
eval64=$(ln $D "void eval64("); hit=$(ln $D "^struct PixelHit"); tmask=$(ln $D "unsigned tile_mask("); ecand=$(ln $D "void eval_candidates(")
rpix=$(ln $D "bool resolve_pixel("); rfull=$(ln $D "void resolve_pixel_full("); foot=$(ln $D "^// -* warp footprints"); hot=$(ln $D "^// -* the hot loop")
prep=$(ln $D "window preparation"); epi=$(ln $D "per-pixel epilogue values"); tau=$(ln $D "float tau32(")
sp=$(ln $F "void store_pixel("); p0=$(ln $F "---------------- P0"); p1=$(ln $F "---------------- P1"); p2=$(ln $F "---------------- P2")
p3=$(ln $F "---------------- P3"); p4=$(ln $F "---------------- P4"); pend=$(ln $F "k_scan_probe")
G="packed asm (scan inner loop):wfot_device.cuh:42-60,tau32:wfot_device.cuh:$tau-$((tau+3)),eval64:wfot_device.cuh:$eval64-$((hit-1)),tile_mask (FP32 re-evaluation):wfot_device.cuh:$((tmask-1))-$((ecand-2)),eval_candidates:wfot_device.cuh:$((ecand-1))-$((rpix-2)),resolve_pixel:wfot_device.cuh:$((rpix-1))-$((rfull-1)),resolve_warp/full:wfot_device.cuh:$rfull-$((foot-1)),footmap/lane_block:wfot_device.cuh:$foot-$((hot-1)),scan_block:wfot_device.cuh:$hot-$((prep-1)),prep_window:wfot_device.cuh:$prep-$((epi-1)),pixel_values:wfot_device.cuh:$epi-$((epi+60)),store_pixel:wfot_fused.cu:$sp-$((sp+14)),fused P0:wfot_fused.cu:$((p0-12))-$((p1-1)),fused P1 control:wfot_fused.cu:$p1-$((p2-1)),fused P2 marginals:wfot_fused.cu:$p2-$((p3-1)),fused P3:wfot_fused.cu:$p3-$((p4-1)),fused P4 gradient:wfot_fused.cu:$p4-$((pend-1)),block_ot1d:wfot_ot.cuh:1-400"
cp gpurun_out/bench_final.json profiles/r01_bench.json
cp gpurun_out/launches_final.csv profiles/r01_bench_launches.csv
python scripts/ncu_launch_summary.py gpurun_out/launches_final.csv 'launch list of `python bench.py --steps 2 --warmup 3 --cpu-sample 0 --secondary 0` under ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares)' > profiles/r01_bench_launches_summary.txt
for cfg in cfg5 cfg4; do
  if [ $cfg = cfg5 ]; then rep=gpurun_out/prof_fused_$TAG.ncu-rep; out=profiles/r01_k_misfit_grad_ncu_summary.txt; what="cfg5 shape (nt=1024, 256x256 grid), 296 windows (1 per resident CTA)"; cmd="python scripts/prof_fp.py 296 fused cfg5"
  else rep=gpurun_out/prof_fused_cfg4_$TAG.ncu-rep; out=profiles/r01_k_misfit_grad_cfg4_ncu_summary.txt; what="cfg4 shape (nt=61, 79x61 grid), 9472 windows"; cmd="python scripts/prof_fp.py 9472 fused cfg4"; fi
  ncu -i $rep --page source --csv --print-source cuda,sass > gpurun_out/src_$cfg.csv 2>/dev/null
  { echo "ncu --set full --clock-control none --import-source on, k_misfit_grad, $what; $NOTE"
    echo "command: $cmd   (report: $rep, not tracked)"; echo
    python scripts/ncu_raw.py $rep 2>/dev/null; echo
    echo "warp-state samples by source region (scripts/ncu_line_summary.py):"
    NCU_GROUPS="$G" python scripts/ncu_line_summary.py gpurun_out/src_$cfg.csv 14; } > $out
done
python - <<'PY'
import json, re
s = open('profiles/r01_k_misfit_grad_ncu_summary.txt').read()
def val(name):
    m = re.search(name + r"\s+(\w+)\s+([\d.]+)", s); u, v = m.group(1), float(m.group(2))
    return v * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[u]
d = json.load(open('profiles/traffic.json'))
d['dram_bytes_per_window'] = (val("dram__bytes_read.sum") + val("dram__bytes_write.sum")) / 296
json.dump(d, open('profiles/traffic.json', 'w'), indent=1)
print("traffic per window", d['dram_bytes_per_window'])
PY
