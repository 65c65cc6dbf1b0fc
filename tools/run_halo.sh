# products-shape layer, halo exchange variants: N = $1, sweep spec in $2 ("mode:copy_streams:overlap:remote_kernel,...")
N=${1:-2}
exec > gpurun_out/halo_${N}.log 2>&1
R="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== products strong N=$N sweep $2"
SGRACE_HALO_SWEEP=$2 $R bench.py --gpus $N --workload products --steps 10 --warmup 3 2>&1 | grep -E '^\{|sweep|Error|error' | cut -c1-700
