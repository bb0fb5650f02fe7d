mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,power.draw --format=csv > gpurun_out/r02b_box.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r02b_gputests.txt 2>&1
timeout 200 python tools/rn_pass_sweep.py 512 3 > gpurun_out/r02b_rn_sweep.txt 2>&1
( time timeout 600 python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err ) > gpurun_out/r02b_bench_time.txt 2>&1
tail -3 gpurun_out/r02b_gputests.txt; cat gpurun_out/r02b_rn_sweep.txt | tail -12; cat gpurun_out/r02b_bench_time.txt; head -c 600 gpurun_out/r02b_bench.json
