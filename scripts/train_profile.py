"""Where the student's training forward + backward spends its GPU time at full size (VC2 UNet, 1x4x16x40x64): per-family and
per-shape CUDA-event timings of one eager forward and one eager backward (ops.start_profile), or — with `ncu` — the same step
bracketed by cudaProfilerStart/Stop for launch lists / single-kernel captures."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from t2v_turbo_b200 import ops  # noqa: E402
from t2v_turbo_b200.configs import VC2_UNET  # noqa: E402
from t2v_turbo_b200.train_unet import StudentUNet  # noqa: E402
from t2v_turbo_b200.unet import UNetModel  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "families"
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
torch.manual_seed(0)
with torch.device(dev):
    base = UNetModel(**{**VC2_UNET, "time_cond_proj_dim": 256})
with torch.no_grad():
    for prm in base.parameters():
        if prm.dim() > 1 and float(prm.abs().max()) == 0.0:
            prm.normal_(0, 0.02)
s = StudentUNet(base.eval(), r=64, dropout_p=0.1).train()
with torch.no_grad():
    for i in range(0, len(s.arena.shapes), 2):
        s.arena.param(i).normal_(0, 0.02)
s.pack()
x = torch.randn(1, 4, 16, 40, 64, device=dev)
ctx = torch.randn(1, 77, 1024, device=dev)
w_emb = torch.randn(1, 256, device=dev)
ts = torch.tensor([499], device=dev)
d_out = torch.randn(1, 4, 16, 40, 64, device=dev)


def fwd():
    return s(x, ts, context=ctx, fps=16, timestep_cond=w_emb)


fwd()
s.backward(d_out)
torch.cuda.synchronize()
if mode == "ncu":
    torch.cuda.cudart().cudaProfilerStart()
    fwd()
    s.backward(d_out)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("profiled one student forward + backward")
    sys.exit(0)
for phase, fn in (("forward", fwd), ("backward", lambda: s.backward(d_out))):
    for by_tag in (False,):
        ops.start_profile()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        fn()
        t1.record()
        rec = ops.stop_profile(by_tag=by_tag)
        tot = sum(d["ms"] for d in rec.values())
        print(f"== student {phase}: {sum(d['calls'] for d in rec.values())} calls, {tot:.1f} ms of kernels (eager wall {t0.elapsed_time(t1):.1f} ms)")
        for fam, d in sorted(rec.items(), key=lambda kv: -kv[1]["ms"])[:14]:
            tf = d["flops"] / (d["ms"] * 1e-3) / 1e12 if d["ms"] > 0 and d["flops"] else 0.0
            print(f"   {fam:28s} {d['calls']:6d} calls {d['ms']:8.2f} ms  {tf:7.1f} TF/s")
    if phase == "forward":
        pass
# torch-side (non-library) time: dropout masks etc. are torch kernels and are not in the family table
