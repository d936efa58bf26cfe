#!/bin/bash
# A/B of libofdmgan side builds on the primary workload: tools/gpu_ab.sh <variant> ...
#   "default" = lib/libofdmgan.so, "general" = default library with OFDMGAN_SIM_IMPL=general (the one-thread-per-frame kernel)
# Per variant: bench.py timing (CUDA events, no profiler), then one ncu pass of three counters on a 2^21-frame launch
# (warp instructions, duration, issue-slot utilisation).  Lines are appended to gpurun_out/ab.txt.
mkdir -p gpurun_out
for v in "$@"; do
  unset OFDMGAN_LIB OFDMGAN_SIM_IMPL
  if [ "$v" = general ]; then export OFDMGAN_SIM_IMPL=general; elif [ "$v" != default ]; then export OFDMGAN_LIB=$PWD/ofdm-gan-sr_b200/lib/libofdmgan_$v.so; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --skip-also --skip-cpu > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  rc=$?
  if [ $rc = 0 ] && [ -z "$AB_NO_NCU" ]; then
    timeout 300 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_sim -s 3 -c 1 --csv --log-file gpurun_out/ab_$v.ncu.csv python bench.py --steps 1 --warmup 3 --skip-also --skip-cpu --frames-per-gpu 2097152 > /dev/null 2>&1
  fi
  python - "$v" <<'PY' >> gpurun_out/ab.txt
import csv, json, sys
v = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/ab_%s.json" % v).read().strip().split("\n")[-1])
    line = "%-10s ms/2^24=%.4f frames/s=%.4g frac=%.4f" % (v, d["ms_per_step"], d["value"], d.get("roofline", {}).get("frac") or 0)
except Exception as e:
    line = "%-10s FAILED %s %s" % (v, e, open("gpurun_out/ab_%s.err" % v).read()[-400:].replace("\n", " | "))
try:
    rows = [r for r in csv.reader(open("gpurun_out/ab_%s.ncu.csv" % v)) if len(r) > 5 and r[0].isdigit()]
    m = {r[-3]: float(r[-1].replace(",", "")) for r in rows}
    line += "  [ncu 2^21: instr/frame=%.0f issue=%.1f%% us=%.1f]" % (m["smsp__inst_executed.sum"] * 32 / 2097152, m["smsp__issue_active.avg.pct_of_peak_sustained_active"], m["gpu__time_duration.sum"] / 1e3 if m["gpu__time_duration.sum"] > 1e4 else m["gpu__time_duration.sum"])
except Exception as e:
    line += "  [ncu: %s]" % e
print(line)
PY
done
cat gpurun_out/ab.txt
