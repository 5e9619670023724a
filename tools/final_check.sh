cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
timeout 900 python bench.py > $O/check_n1.json 2> $O/check_n1.err; echo "n1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 > $O/check_n2.json 2> $O/check_n2.err; echo "n2 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/check_ref_n2.json 2> $O/check_ref_n2.err; echo "ref rc=$?"
python - <<'P'
import json
for f in ("check_n1","check_n2","check_ref_n2"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        print(f, d.get("impl","ours"), d["n_gpus"], round(d["value"],4), "e2e", round(d["e2e"]["value"],4), d.get("gpu_launches"), (d.get("cpu_baseline") or {}).get("cores"), d["config"]["workload"][:60])
    except Exception as e:
        print(f,"FAILED",e)
P
