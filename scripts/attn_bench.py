"""Attention micro-benchmark (CUDA-graph timed): the UNet's spatial self-/cross-attention shapes."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from t2v_turbo_b200 import ops
dev = "cuda"
CASES = {"self_l0": (16, 2560, 2560, 5, 1), "self_l1": (16, 640, 640, 10, 1), "self_l2": (16, 160, 160, 20, 1),
         "cross_l0": (16, 2560, 77, 5, 16), "cross_l1": (16, 640, 77, 10, 16)}
for name in (sys.argv[1:] or list(CASES)):
    b, lq, lk, heads, div = CASES[name]
    inner = heads * 64
    q = torch.randn(b, lq, inner, device=dev).bfloat16()
    k = torch.randn(b // div, lk, inner, device=dev).bfloat16()
    v = torch.randn(b // div, lk, inner, device=dev).bfloat16()
    out = torch.empty_like(q)
    fn = lambda: ops.attention(q, k, v, heads=heads, scale=0.125, kv_batch_div=div, out=out)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(10): fn()
    graph.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); graph.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 * 1e3
    fl = 4 * b * heads * lq * lk * 64
    print(f"{name:10s} {us:8.1f} us {fl / us / 1e6:8.1f} TF/s", flush=True)
