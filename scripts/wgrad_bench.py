"""Micro-benchmark of t2v_wgrad on the LoRA weight-gradient shapes of the VC2 UNet (CUDA-graph timed)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from t2v_turbo_b200 import ops
dev = "cuda"
CASES = {"lin320_down": ((40960,), 320, None), "lin1280_down": ((2560,), 1280, None), "conv320_down": ((16, 40, 64), 320, "3x3"),
         "tconv320_down": ((1, 16, 2560), 320, "t3"), "geglu_up": ((40960,), 2560, None)}
for name, (pts, c, taps) in CASES.items():
    a = torch.randn(*pts, c, device=dev).bfloat16()
    b = torch.randn(*pts, 64, device=dev).bfloat16()
    tl = None if taps is None else (ops._TAPS_3X3 if taps == "3x3" else ops._TAPS_T3)
    nt = 1 if tl is None else len(tl)
    out = torch.zeros(64, c, nt, device=dev)
    fn = lambda: ops.wgrad(a, b, out, taps=tl, out_strides=(c * nt, nt, 1))
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 * 1e3
    n = 1
    for v in pts: n *= v
    fl = 2 * n * c * 64 * nt
    by = n * (c + 64) * 2 * (1 if tl is None else 1)
    print(f"{name:14s} {us:8.1f} us {fl / us / 1e6:8.1f} TF/s  operand bytes {by / 1e6:7.1f} MB -> {by / us / 1e3:7.1f} GB/s", flush=True)
