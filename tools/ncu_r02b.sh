#!/bin/bash
# Round-2 evidence, second pass (after the patchify / embed rewrite, the patch-mode convolutions and the L = 257 attention):
# ncu --set full of (1) every non-contraction kernel of the query path, (2) one ViT-B/16 block at 512 images,
# (3) the whole-row attention kernel at L = 256 / 257 (ViT-L/14: extra-key mode + the helper warp), (4) the first
# convolution GEMMs of a warm RN50x16 pass; then compute-sanitizer memcheck over the tests that touch the new paths.
set -x
K='patchify_kernel|embed_ln_pre_kernel|layernorm_kernel|l2norm_kernel|ln_f16|proto_softmax_kernel|adapter_conv_kernel|prototypes_kernel|resample_h_kernel|resample_v_norm_kernel|attention_tail_rows_kernel'
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$K" -s 15 -c 15 \
    -o gpurun_out/r02b_rowops python tools/rowops_driver.py 3 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:gemm_tn_kernel|attention6" -s 10 -c 5 \
    -o gpurun_out/r02b_block python tools/block_driver.py 512 4 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:attention6" -s 4 -c 1 \
    -o gpurun_out/r02b_attn257 python tools/gpu_probe.py --case attn_257_big 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:attention6" -s 4 -c 1 \
    -o gpurun_out/r02b_attn256 python tools/gpu_probe.py --case attn_256_big 2>&1 | tail -2
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:gemm_tn_kernel" -s 131 -c 7 \
    -o gpurun_out/r02b_rnconv python tools/rn_driver.py 64 1 2>&1 | tail -2
for r in rowops block attn257 attn256 rnconv; do
  python tools/ncu_table.py gpurun_out/r02b_$r.ncu-rep > gpurun_out/r02b_ncu_${r}_summary.txt 2>&1
done
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -p no:cacheprovider \
    -k "test_attention or attention_rows or rn_tiny or rn_small or RN50 or (last_block and small) or test_linear" \
    > gpurun_out/r02b_sanitizer_memcheck.txt 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02b_sanitizer_memcheck.txt
tail -5 gpurun_out/r02b_sanitizer_memcheck.txt
