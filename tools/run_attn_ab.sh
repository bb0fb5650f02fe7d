echo "== tests default"; timeout 400 python -m pytest tests -m gpu -x -q -k "attention" 2>&1 | tail -3
echo "== tests poly=4 pp=0"; PC_ATTN6_PINGPONG=0 PC_ATTN6_POLY=4 timeout 400 python -m pytest tests -m gpu -x -q -k "attention" 2>&1 | tail -3
for pp in 1 0; do for np in 0 1 2 3; do echo "== pp=$pp poly=$np"; PC_ATTN6_PINGPONG=$pp PC_ATTN6_POLY=$np timeout 100 python tools/gpu_probe.py --case attn_big 2>&1 | grep -E "sustained" | cut -c1-200; done; done
echo "== others"; timeout 200 python tools/gpu_probe.py --timeout 60 --only attn_text_big,attn_50 2>&1 | grep -E "sustained|TIMEOUT|rc=" | cut -c1-170
echo "== trace default"; PC_ATTN_TRACE=1 timeout 60 python tools/gpu_probe.py --case attn_big 2>&1 | grep -v "^$" | tail -19
