cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
nvidia-smi topo -m > $O/topo2.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -x -q -k "multi or one_process or shard or pixel or sort" > $O/t6.log 2>&1; tail -5 $O/t6.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu > $O/b6_n1.json 2> $O/b6_n1.err; echo n1 rc=$?
for red in p2p nccl; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 --reduce $red --no-cpu > $O/b6_n2_$red.json 2> $O/b6_n2_$red.err; echo n2 $red rc=$?
done
python - <<'P'
import json
for f in ("b6_n1", "b6_n2_p2p", "b6_n2_nccl"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), round(d["e2e"]["ms_per_step"], 4), d["windows_ms"]["e2e"])
    except Exception as e:
        print(f, "FAILED", e)
P
tail -3 $O/b6_n2_p2p.err
