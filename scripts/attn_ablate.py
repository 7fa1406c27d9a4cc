"""Two-tile attention kernel: time the T2V_ATTN_V2_VARIANT configurations of attn_fwd2.cu on (16, 2560, 2560, 5).

p<N> / q<N>: N of 16 column pairs per chunk on the FMA pipe (clamped / saturating-FMA range reduction); a1 / a2: ablations
with wrong results (no exponentials / + no per-tile barrier) that bound what the synchronisation structure alone costs.
Prints the error against fp32 torch attention for the variants that compute the real thing.
"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["T2V_ATTN_V2"] = "1"
from t2v_turbo_b200 import ops
dev = "cuda"
b, lq, lk, heads = 16, 2560, 2560, 5
inner = heads * 64
torch.manual_seed(0)
q = torch.randn(b, lq, inner, device=dev).bfloat16()
k = torch.randn(b, lk, inner, device=dev).bfloat16()
v = torch.randn(b, lk, inner, device=dev).bfloat16()
out = torch.empty_like(q)
def ref():
    qf, kf, vf = (t[:2].float().view(2, -1, heads, 64).transpose(1, 2) for t in (q, k, v))
    return torch.softmax(qf @ kf.transpose(-1, -2) * 0.125, -1).matmul(vf).transpose(1, 2).reshape(2, lq, inner)
r = ref()
for var in (sys.argv[1:] or ["", "p0", "p4", "p7", "p16", "q5", "q7", "q9", "a1", "a2", "v1"]):
    if var == "v1":
        os.environ["T2V_ATTN_V2"] = "0"
    os.environ["T2V_ATTN_V2_VARIANT"] = var
    fn = lambda: ops.attention(q, k, v, heads=heads, scale=0.125, kv_batch_div=1, out=out)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    err = ((out[:2].float() - r).norm() / r.norm()).item()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(10): fn()
    graph.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); graph.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 10 * 1e3)
    fl = 4 * b * heads * lq * lk * 64
    print(f"variant {var or 'default':8s} {best:8.1f} us {fl / best / 1e6:8.1f} TF/s  rel-L2 vs fp32 {err:.3e}", flush=True)
