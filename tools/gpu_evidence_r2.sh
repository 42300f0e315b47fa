#!/bin/bash
# Round-2 evidence on one B200 (run under gpurun from the repo root): smoke, the bench line, a BLS12-381 bench line,
# the ncu launch list of one proof, and one `ncu --set full --import-source on` capture of the two dominant kernels.
# Everything lands in gpurun_out/; the summaries judged are copied to profiles/ afterwards (tools/ncu_summary.py,
# tools/ncu_stalls_by_line.py).
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_r2.log 2>&1
tail -2 gpurun_out/smoke_r2.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err
tail -c 600 gpurun_out/bench_r2_final.json
timeout 600 python bench.py --curve BLS12_381 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_bls12381.json 2> gpurun_out/bench_r2_bls12381.err
tail -c 300 gpurun_out/bench_r2_bls12381.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r2_final.csv \
    python bench.py --steps 1 --warmup 1 --inflight 1 --no-cpu-baseline > gpurun_out/launches_r2_final.log 2>&1
wc -l gpurun_out/launches_r2_final.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_ntt_pass8|k_msm_accumulate' -s 24 -c 6 \
    -f -o gpurun_out/ncu_r2_top python tools/ntt_time.py 20 > gpurun_out/ncu_r2_top.log 2>&1
ls -la gpurun_out/ncu_r2_top.ncu-rep
