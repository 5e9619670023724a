#!/bin/bash
# One multi-GPU box: the headline configuration at N = 1, 2, 4, 8 and BASELINE.json configs[4] (3840x2160, depth 12,
# 4096 spp) at N = 1, 2, 4, 8 with the final images compared on the box.  Outputs (small) in gpurun_out/.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
G=$(nvidia-smi -L | wc -l)
nvidia-smi topo -m > $O/topo.txt 2>&1
run() {  # n, extra args..., output name last
  local n=$1; shift; local out=${@: -1}; set -- "${@:1:$(($#-1))}"
  if [ "$n" = 1 ]; then timeout 600 python bench.py --gpus 1 "$@" > $O/$out.json 2> $O/$out.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n "$@" > $O/$out.json 2> $O/$out.err; fi
  echo "$out rc=$?"
}
for n in 1 2 4 8; do [ $n -le $G ] && run $n --steps 100 --warmup 8 --no-cpu --no-extras scale_n$n; done
[ 8 -le $G ] && run 8 --steps 100 --warmup 8 --no-cpu --no-extras --reduce nccl scale_n8_nccl
for n in 8 4 2 1; do [ $n -le $G ] && run $n --config 5 --spp 4096 --save-image /tmp/c5_n$n.npy --no-cpu --no-extras c5_n$n; done
python - <<'P'
import json, os, numpy as np
def psnr(a, b):
    mse = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    return 99.0 if mse == 0 else 10 * np.log10(1.0 / mse)
res = {}
for f in sorted(os.listdir("gpurun_out")):
    if f.startswith(("scale_n", "c5_n")) and f.endswith(".json"):
        try:
            d = json.load(open("gpurun_out/" + f))
            res[f[:-5]] = {k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "spp", "seconds", "ms_per_frame", "ms_per_iteration")}
            res[f[:-5]]["e2e"] = d["e2e"]["value"]
            res[f[:-5]]["e2e_ms_per_step"] = d["e2e"].get("ms_per_step")
        except Exception as e:
            res[f[:-5]] = {"error": str(e)}
ref = "/tmp/c5_n1.npy"
if os.path.exists(ref):
    a = np.clip(np.load(ref), 0, 1)
    for n in (2, 4, 8):
        p = f"/tmp/c5_n{n}.npy"
        if os.path.exists(p):
            b = np.load(p)
            res[f"c5_n{n}"]["psnr_vs_1gpu_db"] = psnr(a, np.clip(b, 0, 1))
            res[f"c5_n{n}"]["bit_identical_to_1gpu"] = bool(np.array_equal(np.load(ref).view(np.uint32), b.view(np.uint32)))
json.dump(res, open("gpurun_out/scale_summary.json", "w"), indent=1)
print(json.dumps(res, indent=1))
P
