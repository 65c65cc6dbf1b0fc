set -e
export SGRACE_STREAM_DRY=1
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:spmm_stream -s 6 -c 1 -f -o gpurun_out/prof_dry python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/prof_ncu.log 2>&1
tail -2 gpurun_out/prof_ncu.log | cut -c1-200
