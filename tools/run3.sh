exec > gpurun_out/run3.log 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/products_breakdown.py 1.0 2>&1 | grep -v Warn | grep -v "ref = "
python bench.py --steps 20 --warmup 3 --no-cpu 2>&1 | python tools/brief.py
