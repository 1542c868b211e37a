#!/bin/bash
# Measures the FP64 peaks on the GPU box and writes profiles/fp64_peak.json (copy it back from gpurun_out/).
# usage (under gpurun): bash tools/run_fp64_peak.sh
set -e
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap \
  --format=csv,noheader -lms 100 > gpurun_out/fp64_peak_clocks.csv &
SMI=$!
./build/fp64_peak_probe > gpurun_out/fp64_peak_raw.json
kill $SMI
python - <<'PY'
import json, statistics
d = json.load(open("gpurun_out/fp64_peak_raw.json"))
rows = [l.strip().split(", ") for l in open("gpurun_out/fp64_peak_clocks.csv") if l.strip()]
mhz = [float(r[0].split()[0]) for r in rows]
pw = [float(r[2].split()[0]) for r in rows]
busy = [m for m, p in zip(mhz, pw) if p > 300]
d["clocks"] = {"samples": len(rows), "sm_mhz_median_under_load": statistics.median(busy) if busy else None,
               "sm_max_mhz": float(rows[0][1].split()[0]) if rows else None, "power_w_max": max(pw) if pw else None,
               "hw_slowdown": any("Active" == r[4] for r in rows), "hw_thermal": any("Active" == r[5] for r in rows),
               "sw_thermal": any("Active" == r[6] for r in rows), "sw_power_cap": any("Active" == r[7] for r in rows)}
d["note"] = "peak measured by this repo's probe (tools/fp64_peak_probe.cu); MEASURED_PEAKS.json has no FP64 entry"
json.dump(d, open("gpurun_out/fp64_peak.json", "w"), indent=1)
print(json.dumps(d))
PY
