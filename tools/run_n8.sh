#!/bin/bash
# default bench at N=8 (weak headline + strong products record), then the chunked halo variant
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench_n8.log 2>&1
SGRACE_HALO_CHUNKS=4 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench_n8c.log 2>&1
python - <<PY
import json
for f in ("gpurun_out/r2_bench_n8.log", "gpurun_out/r2_bench_n8c.log"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"])
        print(json.dumps(d.get("strong"), indent=1)[:3000])
        print("molecule_dp", d.get("molecule_dp", {}).get("value"))
    except Exception as ex:
        print(f, "ERR", ex); print(open(f).read()[-2500:])
PY
