N=${1:-8}
for aff in "" "PC_BENCH_NO_AFFINITY=1"; do
env $aff PC_BENCH_NO_CPU=1 PC_BENCH_CFG5_UTT=0 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 2>gpurun_out/bench${N}.err | grep "^{" > gpurun_out/bench${N}_aff.json; python - <<EOF2
import json
d=json.loads(open("gpurun_out/bench${N}_aff.json").read())
print("affinity '$aff':", "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"], "chosen", d.get("collective"))
EOF2
done
