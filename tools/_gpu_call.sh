timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/check_slab_general.py > gpurun_out/slab_general2.log 2>&1; echo "rc=$?" >> gpurun_out/slab_general2.log
grep -v "^W\|^\[W" gpurun_out/slab_general2.log | tail -30
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench2.log 2> gpurun_out/bench2.err; echo "bench2 rc=$?"
python - <<'P'
import json
for l in open("gpurun_out/bench2.log"):
    if l.startswith("{"):
        d=json.loads(l); print(d["value"], d["ms_per_step"], d["n_gpus"], d["roofline"]["frac"], d["clocks"])
P
