exec > gpurun_out/run3.log 2>&1
python -m pytest tests/test_gpu_molecule.py -m gpu -x -q -k "halo or peer" 2>&1 | grep -E "^E|Error|error|passed|failed" | head -12
