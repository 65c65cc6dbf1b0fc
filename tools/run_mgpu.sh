# multi-GPU runs: N = $1
N=${1:-2}
exec > gpurun_out/mgpu_${N}.log 2>&1
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== products strong N=$N agg_first (halo exchange)"
$R bench.py --gpus $N --workload products --steps 10 --warmup 3 2>&1 | grep '^{' | cut -c1-330
echo "== products strong N=$N reference order (all-gather)"
$R bench.py --gpus $N --workload products --order reference --steps 5 --warmup 3 2>&1 | grep '^{' | cut -c1-230
if [ "$2" = "all" ]; then
echo "== cora_x1024 weak N=$N"
$R bench.py --gpus $N --steps 20 --warmup 3 --no-cpu 2>&1 | grep '^{' | python tools/brief.py
echo "== molecule DP N=$N"
$R bench.py --gpus $N --workload molecule --steps 10 --warmup 3 2>&1 | grep '^{' | cut -c1-260
fi
