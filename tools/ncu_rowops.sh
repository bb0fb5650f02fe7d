#!/bin/bash
# ncu --set full of every non-contraction kernel of the query path (one launch each, warm), then the 4 Linear launches
# and the attention launch of one ViT-B/16 block at the benchmark's micro-batch (M = 512 * 197).
set -x
K='patchify_kernel|embed_ln_pre_kernel|row_stats_kernel|layernorm_kernel|l2norm_kernel|ln_f16|proto_softmax_kernel|adapter_conv_kernel|prototypes_kernel|resample_h_kernel|resample_v_norm'
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$K" -s 15 -c 15 \
    -o gpurun_out/r02_rowops python tools/rowops_driver.py 3 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:gemm_tn_kernel|attention6" -s 10 -c 5 \
    -o gpurun_out/r02_block python tools/block_driver.py 512 4 2>&1 | tail -3
