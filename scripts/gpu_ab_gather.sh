for ov in 0 1; do FMK_GATHER_OVERLAP=$ov timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$ov bench.py --gpus 2 --steps 8 --warmup 3 --no-e2e 2>/dev/null > gpurun_out/ab_$ov.json; python - <<PY
import json
d=json.loads(open("gpurun_out/ab_$ov.json").read().strip().split("\n")[-1]); k=d["roofline"]["all_kernels_ms_per_step"]
print("overlap=$ov", d["ms_per_step"], round(sum(k.values()),3), d["config"]["gather_bytes_per_step"])
PY
done
