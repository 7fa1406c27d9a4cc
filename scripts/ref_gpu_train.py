"""The training step's GPU denominator: the UNMODIFIED reference (oracle/_ref snapshot) student — the VC2 `UNetModel` with LoRA
injected by the reference's own `inject_trainable_lora_extended` (r = 64, dropout 0.1 default, train mode) — forward + backward
under `torch.autocast(bfloat16)` on the same B200, one 16x320x512 sample (1x4x16x40x64 latents, 77x1024 context), random-init
weights, the config's `use_checkpoint` (gradient checkpointing) as the reference trains.  Also the teacher forward (no grad).
    --attn naive | sdpa : as scripts/ref_gpu_bench.py (the reference's xformers call served by torch SDPA)
CUDA-event timed, 2 warm-ups, median of N.  Prints one JSON line."""
import argparse
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--attn", default="sdpa", choices=["naive", "sdpa"])
ap.add_argument("--calls", type=int, default=3)
ap.add_argument("--no-checkpoint", action="store_true")
a = ap.parse_args()
if a.attn == "sdpa":
    sys.path.insert(0, os.path.join(ROOT, "oracle", "shim_xformers"))
from oracle import ref_loader  # noqa: E402
assert ref_loader.enable() is not None, "no copy of the reference (run oracle/build_ref.py where /root/reference exists)"
from bench import ClockSampler  # noqa: E402
from t2v_turbo_b200.configs import VC2_UNET  # noqa: E402
from lvdm.modules.networks.openaimodel3d import UNetModel  # noqa: E402
from utils.lora import inject_trainable_lora_extended  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
cfg = {**VC2_UNET, "time_cond_proj_dim": 256, "use_checkpoint": not a.no_checkpoint}
g = torch.Generator(device=dev).manual_seed(0)
with torch.device(dev):
    unet = UNetModel(**cfg)
with torch.no_grad():
    for name, p in unet.named_parameters():
        if p.dim() >= 2:
            p.copy_(torch.randn(p.shape, generator=g, device=dev) * (0.6 / p[0].numel() ** 0.5))
unet.requires_grad_(False)
inject_trainable_lora_extended(unet, target_replace_module={"UNetModel"}, r=64)
unet.to(dev).train()
lora_params = [p for n, p in unet.named_parameters() if "lora_" in n]
n_lora = sum(p.numel() for p in lora_params)
with torch.no_grad():
    for p in lora_params:
        p.requires_grad_(True)
        if float(p.abs().max()) == 0.0:
            p.normal_(0, 0.02)
x = torch.randn(1, 4, 16, 40, 64, device=dev, generator=g)
ctx = torch.randn(1, 77, 1024, device=dev, generator=g)
w_emb = torch.randn(1, 256, device=dev, generator=g)
ts = torch.tensor([499], device=dev)
d_out = torch.randn(1, 4, 16, 40, 64, device=dev, generator=g)


def fwd_bwd():
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = unet(x, ts, context=ctx, fps=16, timestep_cond=w_emb)
    e[1].record()
    (y.float() * d_out).sum().backward()
    e[2].record()
    torch.cuda.synchronize()
    for p in lora_params:
        p.grad = None
    return e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])


for _ in range(2):
    fwd_bwd()
sampler = ClockSampler(0)
sampler.start()
rec = [fwd_bwd() for _ in range(a.calls)]
clocks = sampler.stop()
f_ms, b_ms = statistics.median(r[0] for r in rec), statistics.median(r[1] for r in rec)
print(json.dumps(dict(what="unmodified reference student (LoRA r=64 injected by utils/lora.py, train mode, autocast bf16): forward + backward of "
                           "one 16x320x512 sample on this B200", attn=a.attn, use_checkpoint=not a.no_checkpoint, lora_params=n_lora,
                      forward_ms=f_ms, backward_ms=b_ms, fwd_bwd_ms=f_ms + b_ms, peak_mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30,
                      clocks=clocks)))
