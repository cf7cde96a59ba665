# Round-2 evidence capture on one B200 (TAG=r2): bench lines for every config, the ncu launch list of the headline workload,
# ncu full captures of round 1 (round1_tma_kernel) and rounds 2-3 (round_tc_kernel) at config 3 and config 4 (d = 4, the
# 255-register build), and of the GKR initialisers.  Outputs under gpurun_out/.  The resident kernel talks to the host while it
# runs, so it cannot be replayed by ncu: the full captures run with SC_NO_RESIDENT=1 (the large rounds are unaffected).
set -x
TAG=${TAG:-r2c}
O=gpurun_out
export SC_RES_TIMEOUT_S=5
python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
for c in 1 2 4 5; do python bench.py --config $c --steps 10 --warmup 3 > $O/${TAG}_bench_cfg$c.json 2>> $O/${TAG}_bench.err; done
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2>> $O/${TAG}_bench.err
# ncu makes every launch synchronous, and the resident kernel waits for constants the host writes AFTER the launch call
# returns: under ncu all rounds are launched one by one (SC_NO_RESIDENT=1); rounds 1-7 of nv = 24 are the same kernels either way
export SC_NO_RESIDENT=1
# launch list: 4 proofs (8 eager round-1 chunk launches at prover_init, then 24 launches per proof)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches.csv python tools/prof_run.py 3 4 > $O/${TAG}_launches.log 2>&1
# full captures
# config 3 runs the contraction kernels (gemm_sum.cuh): 8 round-1 chunk launches at prover_init, proof 1 = 6 fold launches (its round 1
# was summed behind the upload), proofs 2.. = round 1 + 6 fold launches -> skip 21: rounds 1, 2, 3 of the third proof
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_round1|gemm_fold' -s 21 -c 3 -o $O/${TAG}_full_cfg3 python tools/prof_run.py 3 4 > $O/${TAG}_ncu_cfg3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'round1_tma|round_tc' -s 24 -c 2 -o $O/${TAG}_full_cfg4 python tools/prof_run.py 4 4 > $O/${TAG}_ncu_cfg4.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'gkr_phase|lanes_normalise|eq_halves|eq_outer' -s 18 -c 6 -o $O/${TAG}_full_gkr python tools/prof_run.py 5 4 > $O/${TAG}_ncu_gkr.log 2>&1
unset SC_NO_RESIDENT
for f in cfg3 cfg4 gkr; do
  ncu -i $O/${TAG}_full_$f.ncu-rep --page raw --csv > $O/${TAG}_full_${f}_raw.csv 2>/dev/null
done
ncu -i $O/${TAG}_full_cfg3.ncu-rep --page source --csv > $O/${TAG}_full_cfg3_source.csv 2>/dev/null
ncu -i $O/${TAG}_full_cfg4.ncu-rep --page source --csv > $O/${TAG}_full_cfg4_source.csv 2>/dev/null
# gpurun_out/ is merged back only below 64 MiB: keep the CSV exports, drop the reports
rm -f $O/*.ncu-rep
gzip -f $O/${TAG}_full_cfg3_source.csv $O/${TAG}_full_cfg4_source.csv
ls -la $O | head -40
