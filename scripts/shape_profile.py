"""Eager per-shape CUDA-event profile of one UNet forward and the VAE decode."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import FRAMES, HEIGHT, WIDTH, build_pipeline  # noqa: E402
from t2v_turbo_b200 import ops  # noqa: E402

BS = int(sys.argv[1]) if len(sys.argv) > 1 else 1       # videos per call
TOP = int(sys.argv[2]) if len(sys.argv) > 2 else 70
dev = torch.device("cuda", 0)
pipe = build_pipeline(dev, use_graph=False)
gen = torch.Generator(device=dev).manual_seed(1)
pe = torch.randn(BS, 77, 1024, device=dev, dtype=torch.bfloat16, generator=gen)
lat = torch.randn(BS, 4, FRAMES, HEIGHT // 8, WIDTH // 8, device=dev, dtype=torch.bfloat16, generator=gen)
ts = torch.full((BS,), 999, device=dev, dtype=torch.long)
wemb = pipe.get_w_embedding(torch.tensor([7.5]).repeat(BS), 256).to(dev).to(torch.bfloat16)
out = {}
for what in ("unet", "vae"):
    fn = (lambda: pipe.unet(lat, ts, context=pe, fps=16, timestep_cond=wemb)) if what == "unet" else \
         (lambda: pipe.pretrained_t2v.decode_first_stage_2DAE(lat))
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    acc = {}
    reps = 3
    for _ in range(reps):
        ops.start_profile()
        fn()
        for k, v in ops.stop_profile(by_tag=True).items():
            d = acc.setdefault(k, dict(calls=0, ms=0.0, flops=0))
            d["calls"] += v["calls"]; d["ms"] += v["ms"]; d["flops"] += v["flops"]
    for v in acc.values():
        v["calls"] //= reps; v["ms"] /= reps; v["flops"] //= reps
    out[what] = acc
    tot = sum(v["ms"] for v in acc.values())
    print(f"== {what} bs={BS}: {tot:.2f} ms (sum of per-call event times)")
    for k, v in sorted(acc.items(), key=lambda kv: -kv[1]["ms"])[:TOP]:
        tf = v["flops"] / max(v["ms"], 1e-9) / 1e9
        print(f"{v['ms']:8.3f} ms x{v['calls']:3d} {tf:7.1f} TF/s  {k}")
json.dump(out, open(os.path.join("gpurun_out", f"shape_profile_bs{BS}.json"), "w"), indent=1)
