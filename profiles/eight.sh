N=${1:-8}
for cfg in "X=1" "NCCL_ALGO=NVLS" "NCCL_ALGO=Ring NCCL_PROTO=LL" "NCCL_ALGO=Tree NCCL_PROTO=LL" "NCCL_ALGO=Ring NCCL_PROTO=LL128"; do
echo "== $cfg"
env $cfg timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 profiles/exp_peer.py 2>&1 | grep "mix 16 nccl\|mix 64 nccl\|mix 16 peer us\|rror" | head -8
done
