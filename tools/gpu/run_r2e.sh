#!/bin/bash
# round 2, call E: full GPU suite, bench line with workloads, ncu captures of the fused Lanczos kernel, f4 study
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2e_rc.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2e_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2e_rc.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?" >> gpurun_out/r2e_rc.txt
timeout 600 python tools/study_online_eigs.py > gpurun_out/r2e_study.md 2> gpurun_out/r2e_study.err; echo "study rc=$?" >> gpurun_out/r2e_rc.txt
REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'lanczos_ritz' -c 4 -o gpurun_out/r2e_prof_fused -f python tools/prof_lanczos.py > gpurun_out/r2e_ncu.log 2>&1; echo "ncu rc=$?" >> gpurun_out/r2e_rc.txt
tail -5 gpurun_out/r2e_tests.log; cat gpurun_out/r2e_rc.txt; cat gpurun_out/r2e_study.md; tail -3 gpurun_out/r2e_study.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2e_bench.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'])
for k,v in d['workloads'].items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ('ms','graphs_per_s','frac_hbm','molecules_per_s','alg_GBs')})
PY
