# Round-1 (second session) evidence capture on one B200: bench line, ncu launch list of the same command, ncu full capture
# of round 1 (round1_tma_kernel) and rounds 2-3 (round_tc_kernel) of the timed proof.  Outputs under gpurun_out/.
set -x
TAG=${TAG:-r1b}
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 600 gpurun_out/${TAG}_bench.json
# launch list: 3 warm-up proofs + per-round-timing proofs precede; list two whole proofs of the timed region
ncu --metrics gpu__time_duration.sum --clock-control none -s 72 -c 48 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'round1_tma|round_tc' -s 30 -c 3 -o gpurun_out/${TAG}_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/
