"""Micro-benchmark of gemm_tc on the UNet's characteristic shapes (CUDA events, L2-warm back-to-back launches).
usage: gemm_bench.py [case ...]   (no args = all); with ncu: -k regex:gemm_tc -c N"""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from t2v_turbo_b200 import ops

BF16 = torch.bfloat16
dev = "cuda"
CASES = {
    # name: (kind, M, K, N, extras)
    "lin320_res": ("linear", 40960, 320, 320, dict(res=True)),
    "lin320": ("linear", 40960, 320, 320, dict()),
    "qkv320": ("linear", 40960, 320, 960, dict()),
    "geglu320": ("linear", 40960, 320, 2560, dict(geglu=True)),
    "ff2_320": ("linear", 40960, 1280, 320, dict(res=True)),
    "lin640_res": ("linear", 10240, 640, 640, dict(res=True)),
    "geglu640": ("linear", 10240, 640, 5120, dict(geglu=True)),
    "lin1280_res": ("linear", 2560, 1280, 1280, dict(res=True)),
    "geglu1280": ("linear", 2560, 1280, 10240, dict(geglu=True)),
    "lin1280_l3": ("linear", 640, 1280, 1280, dict(res=True)),
    "conv320": ("conv", (16, 40, 64, 320), 0, 320, dict(res=True)),
    "conv640": ("conv", (16, 20, 32, 640), 0, 640, dict(res=True)),
    "conv1280": ("conv", (16, 10, 16, 1280), 0, 1280, dict(res=True)),
    "conv1280_l3": ("conv", (16, 5, 8, 1280), 0, 1280, dict(res=True)),
    "tconv320": ("tconv", (1, 16, 2560, 320), 0, 320, dict()),
    "vae512": ("conv", (16, 80, 128, 512), 0, 512, dict(res=True)),
    "vae128": ("conv", (16, 320, 512, 128), 0, 128, dict(res=True)),
    "vae256": ("conv", (16, 160, 256, 256), 0, 256, dict(res=True)),
    "big1280": ("linear", 40960, 1280, 1280, dict(res=True)),
    "tconv1280_l3": ("tconv", (1, 16, 40, 1280), 0, 1280, dict()),
}
CASES.update({
    "qkv640": ("linear", 10240, 640, 1920, dict()),
    "qkv1280": ("linear", 2560, 1280, 3840, dict()),
    "q640": ("linear", 10240, 640, 640, dict()),
    "ff2_640": ("linear", 10240, 2560, 640, dict(res=True)),
    "ff2_1280": ("linear", 2560, 5120, 1280, dict(res=True)),
    "tconv1280": ("tconv", (1, 16, 160, 1280), 0, 1280, dict()),
    "conv1280_cat": ("conv", (16, 10, 16, 2560), 0, 1280, dict(res=True)),
})
BN = int(os.environ.get("BN", "0"))   # force a tile width (experiments)
STATS = int(os.environ.get("STATS", "0"))   # conv / tconv: also accumulate the GroupNorm statistics in the epilogue
names = sys.argv[1:] or list(CASES)
for name in names:
    kind, M, K, N, ex = CASES[name]
    g = torch.Generator(device=dev).manual_seed(0)
    if kind == "linear":
        x = torch.randn(M, K, device=dev, generator=g).to(BF16)
        w = (torch.randn(N, K, device=dev, generator=g) * K ** -0.5).to(BF16)
        b = torch.randn(N, device=dev, generator=g)
        n_out = N // 2 if ex.get("geglu") else N
        res = torch.randn(M, n_out, device=dev, generator=g).to(BF16) if ex.get("res") else None
        out = torch.empty(M, n_out, device=dev, dtype=BF16)
        fn = lambda: ops.linear(x, w, b, residual=res, geglu=bool(ex.get("geglu")), out=out, block_n=BN)
        flops = 2 * M * K * N
    elif kind == "conv":
        n, h, wd, c = M
        x = torch.randn(n, h, wd, c, device=dev, generator=g).to(BF16)
        w = (torch.randn(N, 9 * c, device=dev, generator=g) * (9 * c) ** -0.5).to(BF16)
        b = torch.randn(1, N, device=dev, generator=g)
        res = torch.randn(n, h, wd, N, device=dev, generator=g).to(BF16)
        out = torch.empty(n, h, wd, N, device=dev, dtype=BF16)
        st = torch.zeros(n, N, 2, device=dev) if STATS else None
        fn = lambda: ops.conv3x3(x, w, b, bias_div=n, residual=res, out=out, block_n=BN, stats=st)
        flops = 2 * n * h * wd * 9 * c * N
    else:
        bb, t, hw, c = M
        x = torch.randn(bb, t, hw, c, device=dev, generator=g).to(BF16)
        w = (torch.randn(N, 3 * c, device=dev, generator=g) * (3 * c) ** -0.5).to(BF16)
        b = torch.randn(N, device=dev, generator=g)
        out = torch.empty(bb, t, hw, N, device=dev, dtype=BF16)
        st = torch.zeros(bb * t, N, 2, device=dev) if STATS else None
        fn = lambda: ops.tconv3(x, w, b, out=out, block_n=BN, stats=st)
        flops = 2 * bb * t * hw * 3 * c * N
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    # GPU-only time: 20 back-to-back launches captured in a CUDA graph (no host launch overhead)
    reps = 20
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(reps):
            fn()
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print(f"{name:14s} {us:8.1f} us  {flops / us / 1e6:8.1f} TF/s", flush=True)
