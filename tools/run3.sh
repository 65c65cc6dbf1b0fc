exec > gpurun_out/run3.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu 2>&1 | tee gpurun_out/bench_r1e.log | python tools/brief.py
grep -o '"single_graph": {"layer_us": [0-9.]*, "launches_per_layer": [0-9.]*' gpurun_out/bench_r1e.log
timeout 100 python tools/dbg_fused.py 2>&1 | head -5
