N=${1:-2}
nvidia-smi topo -m 2>&1 | head -12
NCCL_DEBUG=WARN python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 scripts/nccl_probe.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -12
