#!/bin/bash
# One gpurun call: bench lines of the three headline workloads, the ncu launch list of the default
# bench command, and an ncu --set full capture of one training step's kernels.  Output -> gpurun_out/.
#   tools/capture_profiles.sh TAG [bench|all]
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
WHAT=${2:-all}
python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_lstm.json 2> gpurun_out/${TAG}_bench_lstm.err; echo "bench lstm rc=$?"
python bench.py --workload attn_gru_train --steps 10 --warmup 5 > gpurun_out/${TAG}_bench_attn_gru.json 2> gpurun_out/${TAG}_bench_attn_gru.err; echo "bench attn rc=$?"
python bench.py --workload beam3 --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_beam3.json 2> gpurun_out/${TAG}_bench_beam3.err; echo "bench beam rc=$?"
cat gpurun_out/${TAG}_bench_lstm.json gpurun_out/${TAG}_bench_attn_gru.json gpurun_out/${TAG}_bench_beam3.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['workload'][:40], '| value', round(d['value']), '| e2e', round(d['e2e']['value']), '| ms', round(d['ms_per_step'], 3), '| frac', d['roofline']['frac'], '| launches', d['gpu_launches']); print('   ', d['roofline']['kernels_ms'])
"
[ "$WHAT" = "bench" ] && exit 0
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_lstm.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'gemm_tc|rnn_seq_tc|rnn_cluster' -s 52 -c 14 -f -o gpurun_out/${TAG}_lstm_full $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
tail -2 gpurun_out/ncu_full.log
