exec > gpurun_out/run3.log 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --workload molecule --steps 20 --warmup 3 2>&1 | tail -1 | cut -c1-330
./tools/micro/tc_fea_test 227732 64 64 | tail -2
