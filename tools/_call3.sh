mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_preprocess.py -m gpu -x -q 2>&1 | tail -4 ) > gpurun_out/r02d_pp_tests.txt 2>&1
timeout 200 python tools/gpu_probe.py --only preprocess > gpurun_out/r02d_pp_probe.txt 2>&1
K='resample_h_kernel|resample_v_norm_kernel'
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$K" -s 4 -c 2 -o gpurun_out/r02d_pp python tools/rowops_driver.py 3 > gpurun_out/r02d_pp_ncu.log 2>&1
timeout 200 compute-sanitizer --tool memcheck python -m pytest tests/test_preprocess.py -m gpu -x -q > gpurun_out/r02d_pp_memcheck.txt 2>&1
cat gpurun_out/r02d_pp_tests.txt; grep preprocess gpurun_out/r02d_pp_probe.txt; tail -3 gpurun_out/r02d_pp_ncu.log; tail -4 gpurun_out/r02d_pp_memcheck.txt
