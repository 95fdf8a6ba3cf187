#!/bin/bash
# usage (under gpurun --gpus N): tools/run_scale.sh N  -> gpurun_out/r2_bench_nN.json
N=$1
if [ "$N" = "1" ]; then
  python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
fi
tail -2 gpurun_out/r2_bench_n$N.err
python tools/show_bench.py gpurun_out/r2_bench_n$N.json | head -2
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_n$N.json").read().strip().splitlines()[-1])
print("dp_parity", d.get("dp_parity"))
e=d["e2e"]; print("e2e", {k:e[k] for k in e if k not in ("api",)})
print("enc_fwd", (d.get("encoder_fwd") or {}).get("value"), "enc_train", (d.get("encoder_train") or {}).get("value"), "config1", d.get("config1_fwd"))
ig=d.get("in_graph") or {}
print("in_graph step_us", ig.get("step_us"), {k:v["us"] for k,v in (ig.get("kernels") or {}).items()})
PY
