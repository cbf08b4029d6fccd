#!/bin/bash
# round 2, validation: full GPU suite, smoke, bench line with workloads, reference arm (short)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2y_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2y_rc.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2y_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2y_rc.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; echo "bench rc=$?" >> gpurun_out/r2y_rc.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2y_ref.json 2> gpurun_out/r2y_ref.err; echo "ref rc=$?" >> gpurun_out/r2y_rc.txt
tail -5 gpurun_out/r2y_tests.log; tail -3 gpurun_out/r2y_smoke.log; cat gpurun_out/r2y_rc.txt; tail -3 gpurun_out/r2y_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2y_bench.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'], 'roof', d['roofline'].get('frac'), 'launches', d['gpu_launches'], 'clocks', d['clocks'])
for k,v in d['workloads'].items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ('ms','graphs_per_s','frac_hbm','molecules_per_s','alg_GBs','ms_eager','ms_graphed','error','cpu_port')})
r=json.load(open('gpurun_out/r2y_ref.json')); print('ref', r.get('value'), r.get('unit'), r.get('cpu_baseline'))
PY
