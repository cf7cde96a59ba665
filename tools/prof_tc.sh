# ncu full capture of the tensor-core fold kernel: rounds 2 and 3 of the 4th proof (3 warm-up proofs x 9 TC launches)
export SC_DEBUG=1
SKIP=${SKIP:-27}
timeout 300 ncu --set full --clock-control none --import-source on -k regex:round_tc -s $SKIP -c 2 -o gpurun_out/${OUT:-tc_r2} python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_tc.log 2>&1
tail -2 gpurun_out/ncu_tc.log | cut -c1-300
timeout 200 python bench.py --no-cpu-baseline --steps 5 --warmup 3 2>gpurun_out/err.log > gpurun_out/bench_tc.json
python -c "
import json; d=json.loads(open('gpurun_out/bench_tc.json').read().strip().splitlines()[-1]); print(round(d['value']/1e9,2), 'G/s', round(d['ms_per_step'],3), 'ms', d['config']['round_ms'][:10])"
grep round_tc gpurun_out/err.log | head -2
