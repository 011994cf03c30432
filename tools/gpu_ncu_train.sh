#!/bin/bash
# ncu --set full of the train path's kernels that matter (one eager step of tools/train_step_once.py; warm caches).
# Launch skip counts land in the decoder loop (the first launches of a kernel belong to the encoder LSTM / pre-loop).
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --cache-control none --import-source on"
$NCU -k regex:skinny_nn_strip -s 200 -c 4 -o gpurun_out/r2_train_nn -f python tools/train_step_once.py 8 1 > gpurun_out/r2_ncu_train_nn.log 2>&1
$NCU -k regex:skinny_nt_smem -s 300 -c 6 -o gpurun_out/r2_train_nt -f python tools/train_step_once.py 8 1 > gpurun_out/r2_ncu_train_nt.log 2>&1
$NCU -k regex:"attn_step|sgemm_tn_rows" -s 40 -c 4 -o gpurun_out/r2_train_attn -f python tools/train_step_once.py 8 1 > gpurun_out/r2_ncu_train_attn.log 2>&1
for n in nn nt attn; do ncu -i gpurun_out/r2_train_$n.ncu-rep --page raw --csv > gpurun_out/r2_train_${n}_raw.csv 2>/dev/null; done
ls -la gpurun_out/r2_train_*
