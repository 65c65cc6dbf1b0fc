# ncu capture of the two streaming SpMM launches (FEA, ADJ) of one bench step
set -e
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:spmm_stream -s 6 -c 2 -f -o gpurun_out/prof_stream python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/prof_ncu.log 2>&1
tail -3 gpurun_out/prof_ncu.log
