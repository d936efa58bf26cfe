#!/bin/bash
# A/B of critic-kernel side builds: tools/debug/critic_ab.sh <variant> ... ("default" = lib/libofdmgan.so)
mkdir -p gpurun_out
for v in "$@"; do
  unset OFDMGAN_LIB
  if [ "$v" != default ]; then export OFDMGAN_LIB=$PWD/ofdm-gan-sr_b200/lib/libofdmgan_$v.so; fi
  python tools/debug/time_train.py > gpurun_out/cab_$v.txt 2> gpurun_out/cab_$v.err
  ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__cycles_active.max,sm__cycles_active.avg,sm__cycles_elapsed.max --clock-control none -k regex:k_critic2 -s 6 -c 1 --csv --log-file gpurun_out/cab_$v.ncu.csv python tools/run_train_step.py 3 > /dev/null 2>&1
  python - "$v" <<'PY'
import csv, json, sys
v = sys.argv[1]
line = "%-10s %s" % (v, open("gpurun_out/cab_%s.txt" % v).read().strip() or open("gpurun_out/cab_%s.err" % v).read()[-300:])
rows = [r for r in csv.reader(open("gpurun_out/cab_%s.ncu.csv" % v)) if len(r) > 5 and r[0].isdigit()]
m = {r[-3]: float(r[-1].replace(",", "")) for r in rows}
print(line, " ".join("%s=%.4g" % (k.split("__")[1][:28], x) for k, x in m.items()))
PY
done
