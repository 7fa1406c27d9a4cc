"""Context number (not the judged baseline): the reference's own arithmetic — the plain-torch restatement in oracle/,
i.e. the same cuDNN / cuBLAS / ATen ops the unmodified reference modules dispatch to — run in bf16 on the GPU,
eager, naive attention (what the reference executes when xformers is absent, attention.py:128-144)."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from t2v_turbo_b200.configs import VC2_UNET, VC2_VAE_DDCONFIG
from t2v_turbo_b200.unet import UNetModel
from t2v_turbo_b200.vae import AutoencoderKL
from oracle.unet_oracle import unet_forward, guidance_scale_embedding
from oracle.vae_oracle import decode_first_stage_2dae

dev = torch.device("cuda")
with torch.device("meta"):
    shapes = {k: v.shape for k, v in UNetModel(**VC2_UNET).state_dict().items()}
    vshapes = {k: v.shape for k, v in AutoencoderKL(VC2_VAE_DDCONFIG, 4).state_dict().items()}
g = torch.Generator(device=dev).manual_seed(0)
def mk(shapes):
    sd = {}
    for k, shp in shapes.items():
        if len(shp) >= 2:
            fan = 1
            for s in shp[1:]:
                fan *= s
            sd[k] = (torch.randn(tuple(shp), device=dev, generator=g) * (0.6 / fan ** 0.5)).bfloat16()
        elif k.endswith("weight"):
            sd[k] = torch.ones(tuple(shp), device=dev, dtype=torch.bfloat16)
        else:
            sd[k] = torch.zeros(tuple(shp), device=dev, dtype=torch.bfloat16)
    return sd
sd, vsd = mk(shapes), mk(vshapes)
x = torch.randn(1, 4, 16, 40, 64, device=dev, generator=g).bfloat16()
ctx = torch.randn(1, 77, 1024, device=dev, generator=g).bfloat16()
w = guidance_scale_embedding(torch.tensor([7.5]), 256).to(dev).bfloat16()
ts = torch.tensor([999], device=dev)
def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
with torch.no_grad():
    t_unet = timeit(lambda: unet_forward(sd, VC2_UNET, x, ts, ctx, fps=16, timestep_cond=w))
    t_vae = timeit(lambda: decode_first_stage_2dae(vsd, VC2_VAE_DDCONFIG, x), n=2)
pipe_ms = 4 * t_unet + t_vae
print(json.dumps(dict(what="torch eager bf16 (reference arithmetic via oracle restatement) on B200", unet_fwd_ms=t_unet,
                      vae_decode_16f_ms=t_vae, pipeline_4step_ms_est=pipe_ms, frames_per_s_est=16 / (pipe_ms * 1e-3))))
