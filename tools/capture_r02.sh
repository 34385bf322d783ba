#!/bin/bash
# Round-2 profile capture, one gpurun call.  Every ncu pass follows a plain run of the same command that exited 0.
# Reports stay on the box (they exceed what gpurun_out/ carries back); their raw pages come back as CSV.
set -u
mkdir -p gpurun_out
T=r02
# 1. launch list of the default bench command (main workload only)
CMD="python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline --no-gpu-reference"
$CMD > gpurun_out/${T}_plain_lstm.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${T}_launches_lstm.csv $CMD > gpurun_out/${T}_ncu_list_lstm.log 2>&1
echo "list lstm rc=$?"
# 2. ncu --set full of one eager LSTM training step: every GEMM and both recurrent kernels
STEP="python tools/ncu_step.py lstm 256 0 bf16"
$STEP > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gemm_tc|rnn_seq_tc|rnn_cluster" -f -o /tmp/${T}_lstm_full $STEP > gpurun_out/${T}_ncu_full_lstm.log 2>&1
echo "full lstm rc=$?"
ncu -i /tmp/${T}_lstm_full.ncu-rep --page raw --csv > gpurun_out/${T}_lstm_full_raw.csv 2>/dev/null
ls -la /tmp/${T}_lstm_full.ncu-rep
[ $(stat -c %s /tmp/${T}_lstm_full.ncu-rep) -lt 30000000 ] && cp /tmp/${T}_lstm_full.ncu-rep gpurun_out/
# 3. attention-GRU: the big GEMMs, one recurrent step pair and one attention step pair
STEP="python tools/ncu_step.py attn_gru 128 196 bf16"
$STEP > /dev/null 2>&1 && \
ncu --set full --clock-control none --profile-from-start off -k regex:"gemm_tc|rnn_step_x|attn_stream|relayout|hoist|embed_q" -c 64 -f -o /tmp/${T}_attn_full $STEP > gpurun_out/${T}_ncu_full_attn.log 2>&1
echo "full attn rc=$?"
ncu -i /tmp/${T}_attn_full.ncu-rep --page raw --csv > gpurun_out/${T}_attn_full_raw.csv 2>/dev/null
ls -la /tmp/${T}_attn_full.ncu-rep
# 4. beam-3 launch list
CMD="python bench.py --workload beam3 --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-reference"
$CMD > gpurun_out/${T}_plain_beam.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${T}_launches_beam3.csv $CMD > gpurun_out/${T}_ncu_list_beam.log 2>&1
echo "list beam rc=$?"
# 5. ncu --set full of the decode-step kernels (screening GEMM, select, W_hh 3xTF32, gate) of a beam-3 call
STEP="python tools/run_beam_once.py 3 4096"
$STEP > /dev/null 2>&1 && \
ncu --set full --clock-control none -k regex:"gemm_tc2|screen_select|tf32x3|decode_gate" -s 40 -c 8 -f -o /tmp/${T}_beam_full $STEP > gpurun_out/${T}_ncu_full_beam.log 2>&1
echo "full beam rc=$?"
ncu -i /tmp/${T}_beam_full.ncu-rep --page raw --csv > gpurun_out/${T}_beam_full_raw.csv 2>/dev/null
du -sh gpurun_out
