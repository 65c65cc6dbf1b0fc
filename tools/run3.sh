exec > gpurun_out/run3.log 2>&1
python -m pytest tests/test_gpu_configs.py -m gpu -x -q 2>&1 | grep -E "^E|passed|failed|Error" | head -20
