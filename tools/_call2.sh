mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q -k "preprocess or train or augmented or rn or RN or resnet or Resnet or cli" 2>&1 | tail -8 ) > gpurun_out/r02c_gputests.txt 2>&1
timeout 200 python tools/rn_pass_sweep.py 512 3 > gpurun_out/r02c_rn_sweep.txt 2>&1
( time timeout 600 python bench.py > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err ) > gpurun_out/r02c_bench_time.txt 2>&1
tail -4 gpurun_out/r02c_gputests.txt; tail -5 gpurun_out/r02c_rn_sweep.txt; head -c 300 gpurun_out/r02c_bench.json
