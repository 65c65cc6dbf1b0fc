exec > gpurun_out/run3.log 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --workload products --steps 3 --warmup 3 2>&1 | grep '^{' | cut -c1-200
