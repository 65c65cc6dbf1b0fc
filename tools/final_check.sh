# end-of-round verification on one GPU: tests, smoke, both bench arms, the other workloads
exec > gpurun_out/final_check.log 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== bench (default)"; timeout 600 python bench.py 2>&1 | grep '^{' | tee gpurun_out/bench_final.json | python tools/brief.py
echo "== bench --impl reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | grep '^{' | cut -c1-400
echo "== bench molecule"; timeout 300 python bench.py --workload molecule --steps 10 --warmup 3 2>&1 | grep '^{' | cut -c1-300
echo "== bench products (1 GPU)"; timeout 300 python bench.py --workload products --steps 5 --warmup 3 2>&1 | grep '^{' | cut -c1-260
