N=${1:-8}
timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --check 2>&1 | grep "^{" > gpurun_out/check_$N.json; cut -c1-60 gpurun_out/check_$N.json
PC_BENCH_NO_CPU=1 timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 2>gpurun_out/bench${N}.err | grep "^{" > gpurun_out/bench${N}_auto.json; python - <<EOF2
import json
d=json.loads(open("gpurun_out/bench${N}_auto.json").read())
print("auto", "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "chosen", d.get("collective"), "trial", d.get("collective_trial_ms_per_step"), "timeouts", d.get("peer_timeouts"))
print("   cfg5 ms/iter", d["cfg5"]["ms_per_iteration"], d["cfg5"]["stage_ms"], d["cfg5"]["collective"], d["cfg5"]["collective_trial_ms_per_iteration"])
EOF2
