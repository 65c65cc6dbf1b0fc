exec > gpurun_out/run1.log 2>&1
for v in "A=1" "SGRACE_STREAM_PF=1" "SGRACE_STREAM_PF=2" "SGRACE_STREAM_PF=4"; do
  echo "== $v"
  env $v python bench.py --steps 20 --warmup 3 --no-cpu 2>&1 | python tools/brief.py
done
