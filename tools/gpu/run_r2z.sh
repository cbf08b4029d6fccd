#!/bin/bash
# ncu --set full over the round-2 kernels without a capture; only the raw CSV page travels back
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
LNB_NO_GRAPH=1 timeout 1500 ncu --set full --clock-control none \
  -k regex:'batch_prepare|graph_messages|operator_chain|gaussian|tridiag_powers|symmetrize|RowLoadPolicy|lanczos_ritz|batched_gemm|readout|embedding' \
  -c 48 -o /tmp/r2z_prof_others -f python tools/prof_others.py > gpurun_out/r2z_ncu.log 2>&1; echo "ncu rc=$?" > gpurun_out/r2z_rc.txt
ncu -i /tmp/r2z_prof_others.ncu-rep --page raw --csv > gpurun_out/r2z_raw.csv 2> gpurun_out/r2z_raw.err; echo "export rc=$?" >> gpurun_out/r2z_rc.txt
cat gpurun_out/r2z_rc.txt; tail -2 gpurun_out/r2z_ncu.log; ls -la gpurun_out/r2z_raw.csv
