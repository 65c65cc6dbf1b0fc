exec > gpurun_out/run2.log 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --workload molecule --steps 10 --warmup 3 2>&1 | tail -3
python bench.py --workload products --scale 0.25 --steps 3 --warmup 3 2>&1 | tail -3
