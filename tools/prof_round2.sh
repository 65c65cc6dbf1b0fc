# Round-2 profile (one GPU; the reports must stay under 64 MiB together to travel back: source import only for the streaming kernel).  Every ncu run follows a plain run of the same command that exited 0.
#  (1) launch list of the headline step, (2) ncu --set full of the two streaming launches (FEA, ADJ),
#  (3) ncu --set full of the tcgen05 dense FEA at the products shape, (4) launch list + ncu --set full of the
#  quantised full-design kernels (PubMed-shape x32), (5) ncu --set full of the HALF C-simulation kernels
R=r2
B="python bench.py --steps 2 --warmup 3 --no-cpu --headline-only"
$B > gpurun_out/${R}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 12 --csv --log-file gpurun_out/${R}_launches.csv $B > gpurun_out/${R}_ncu1.log 2>&1
$B > gpurun_out/${R}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:spmm_stream -s 6 -c 2 -f -o gpurun_out/${R}_spmm_stream $B > gpurun_out/${R}_ncu2.log 2>&1
make -C tools/micro tc_fea_test > /dev/null 2>&1
./tools/micro/tc_fea_test 2449029 100 256 > gpurun_out/${R}_tc_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:fea_dense_tc -s 2 -c 1 -f -o gpurun_out/${R}_fea_dense_tc ./tools/micro/tc_fea_test 2449029 100 256 > gpurun_out/${R}_ncu3.log 2>&1
python tools/gat_bench.py 32 > gpurun_out/${R}_gat_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${R}_gat_launches.csv python tools/gat_bench.py 32 > gpurun_out/${R}_ncu4.log 2>&1
python tools/gat_bench.py 32 > gpurun_out/${R}_gat_plain2.log 2>&1 &&
ncu --set full --clock-control none -k regex:'gat_aggregate_vec|spmm_stream|fea_q_csr' -c 3 -f -o gpurun_out/${R}_gat python tools/gat_bench.py 32 > gpurun_out/${R}_ncu5.log 2>&1
python tools/half_bench.py > gpurun_out/${R}_half_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:stage_exact -s 4 -c 2 -f -o gpurun_out/${R}_half python tools/half_bench.py > gpurun_out/${R}_ncu6.log 2>&1
tail -2 gpurun_out/${R}_tc_plain.log; tail -3 gpurun_out/${R}_gat_plain.log; tail -2 gpurun_out/${R}_half_plain.log; ls -la gpurun_out/${R}_*.ncu-rep
