exec > gpurun_out/run1.log 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 20 --warmup 3 2>&1 | tee gpurun_out/bench_r1c.log | python tools/brief.py
python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1
