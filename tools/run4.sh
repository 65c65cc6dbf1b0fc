exec > gpurun_out/run4.log 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -v Warning | tail -25
timeout 200 python tools/prep_bench.py 2>&1 | tail -6
