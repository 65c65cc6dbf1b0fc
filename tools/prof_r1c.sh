# Round-1 (c) profile: quantised / GAT kernels and the one-launch small layer.
# (1) plain runs, (2) launch list of the full-design layers, (3) ncu --set full of the GAT / GCN aggregation
# kernels, (4) ncu --set full of the fused small-layer kernel
R=r1c
python tools/gat_bench.py 8 > gpurun_out/${R}_gat_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${R}_gat_launches.csv python tools/gat_bench.py 8 > gpurun_out/${R}_ncu1.log 2>&1
python tools/gat_bench.py 8 > gpurun_out/${R}_gat_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'gat_aggregate_vec|adj_q_gcn_vec|fea_q_csr' -c 3 -f -o gpurun_out/${R}_gat python tools/gat_bench.py 8 > gpurun_out/${R}_ncu2.log 2>&1
python tools/dbg_fused.py > gpurun_out/${R}_small_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fused_small -s 5 -c 1 -f -o gpurun_out/${R}_fused_small python tools/dbg_fused.py > gpurun_out/${R}_ncu3.log 2>&1
tail -3 gpurun_out/${R}_gat_plain.log; tail -2 gpurun_out/${R}_ncu2.log; tail -2 gpurun_out/${R}_ncu3.log
