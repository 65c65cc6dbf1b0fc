exec > gpurun_out/run3.log 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
