exec > gpurun_out/run3.log 2>&1
python -m pytest tests/test_gpu_molecule.py -m gpu -x -q 2>&1 | tail -2
python bench.py --workload molecule --steps 20 --warmup 3 2>&1 | tail -1 | cut -c1-330
