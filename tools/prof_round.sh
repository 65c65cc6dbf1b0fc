# Round profile: (1) plain run, (2) launch list with device times, (3) ncu --set full of the two streaming
# launches, (4) ncu --set full of the tensor-core dense FEA (micro harness, 500k x 100 -> 256)
R=${1:-r1}
python __graft_entry__.py smoke > gpurun_out/${R}_smoke.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${R}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 15 -c 10 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${R}_ncu1.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${R}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:spmm_stream -s 6 -c 2 -f -o gpurun_out/${R}_spmm_stream python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${R}_ncu2.log 2>&1
make -C tools/micro tc_fea_test > /dev/null 2>&1
./tools/micro/tc_fea_test 500000 100 256 > gpurun_out/${R}_tc_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fea_dense_tc -s 2 -c 1 -f -o gpurun_out/${R}_fea_dense_tc ./tools/micro/tc_fea_test 500000 100 256 > gpurun_out/${R}_ncu3.log 2>&1
tail -2 gpurun_out/${R}_smoke.log; tail -2 gpurun_out/${R}_tc_plain.log
