"""GroupNorm micro-benchmark: two-kernel path (mode 1) vs single-kernel cluster path (mode 2), CUDA-graph timed."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from t2v_turbo_b200 import ops
dev = "cuda"
CASES = [(16, 163840, 128), (16, 40960, 512), (128, 2560, 320), (128, 640, 640), (128, 160, 1280), (128, 40960, 256), (8, 40960, 320), (8, 10240, 1280)] if "big" in sys.argv[1:] else [(16, 2560, 320), (16, 2560, 640), (16, 640, 640), (16, 640, 1280), (16, 640, 1920), (16, 160, 1280), (16, 160, 2560), (16, 40, 1280), (16, 40, 2560)]
for n, hw, c in CASES:
    x = torch.randn(n * hw, c, device=dev).bfloat16()
    g = torch.randn(c, device=dev); b = torch.randn(c, device=dev)
    out = torch.empty_like(x)
    res = []
    for mode in ((1,) if "big" in sys.argv[1:] else (1, 2)):
        try:
            fn = lambda: ops.groupnorm(x, g, b, rows_per_sample=hw, eps=1e-5, silu=True, out=out, mode=mode)
            for _ in range(3): fn()
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                for _ in range(20): fn()
            graph.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); graph.replay(); e1.record(); torch.cuda.synchronize()
            res.append(e0.elapsed_time(e1) / 20 * 1e3)
        except RuntimeError as e:
            res.append(float("nan"))
    mb = n * hw * c * 2 / 1e6
    res = res + [float("nan")] * (2 - len(res))
    print(f"n={n} hw={hw} c={c} ({mb:.1f} MB): two-kernel {res[0]:.1f} us ({3 * mb / res[0]:.2f} TB/s of 3 passes), cluster {res[1]:.1f} us", flush=True)
