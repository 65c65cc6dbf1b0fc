exec > gpurun_out/run4.log 2>&1
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | grep -v Warning | tail -5
for k in 1 2 0; do echo "FEAQ_KERNEL=$k"; SGRACE_FEAQ_KERNEL=$k timeout 300 python tools/gat_bench.py 32 2>&1 | head -2; done
