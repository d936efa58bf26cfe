python -m pytest tests -m gpu -x -q -k "critic or train or penalty or step or disc" 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 --skip-cpu > gpurun_out/bench_crit.json 2> gpurun_out/bench_crit.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_crit.json').read().strip().split('\n')[-1])
print(json.dumps(d['summary']))
PY
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__cycles_active.max,sm__cycles_active.avg --clock-control none -k regex:k_critic2 -s 6 -c 1 --csv --log-file gpurun_out/crit.ncu.csv python tools/run_train_step.py 3 > /dev/null 2>&1; grep -E "inst_executed|time_duration|issue_active|cycles_active" gpurun_out/crit.ncu.csv | awk -F'","' '{print $(NF-2), $NF}'
