"""The north star's GPU denominator: the UNMODIFIED reference (oracle/_ref snapshot) — `T2VTurboVC2Pipeline.__call__`
driving the reference `UNetModel`, `T2VTurboScheduler` and frame-by-frame `Decoder` — in bf16 on the same B200, 4 steps,
16x320x512, random-init VC2 weights, in two variants:
    --attn naive : what the reference runs without xformers (einsum logits + softmax, attention.py:102-164)
    --attn sdpa  : the reference's own `efficient_forward` flash path (attention.py:166-240) with its one library call
                   (`xformers.ops.memory_efficient_attention`) served by torch SDPA (oracle/shim_xformers)
CUDA-event timed, 2 warm-ups, median of N calls, clocks sampled during the timed region.  Prints one JSON line."""
import argparse
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--attn", default="naive", choices=["naive", "sdpa"])
ap.add_argument("--calls", type=int, default=5)
ap.add_argument("--batch", type=int, default=1)
a = ap.parse_args()
if a.attn == "sdpa":
    sys.path.insert(0, os.path.join(ROOT, "oracle", "shim_xformers"))
from oracle import ref_loader  # noqa: E402
assert ref_loader.enable() is not None, "no copy of the reference (run oracle/build_ref.py where /root/reference exists)"
from bench import ClockSampler, FRAMES, HEIGHT, WIDTH, STEPS  # noqa: E402
from t2v_turbo_b200.configs import VC2_UNET, VC2_VAE_DDCONFIG  # noqa: E402
from lvdm.modules.networks.openaimodel3d import UNetModel  # noqa: E402
from lvdm.modules.networks.ae_modules import Decoder  # noqa: E402
import lvdm.modules.attention as ratt  # noqa: E402
from pipeline.t2v_turbo_vc2_pipeline import T2VTurboVC2Pipeline  # noqa: E402
from scheduler.t2v_turbo_scheduler import T2VTurboScheduler  # noqa: E402

assert ratt.XFORMERS_IS_AVAILBLE == (a.attn == "sdpa")
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
with torch.device("meta"):
    unet, dec = UNetModel(**VC2_UNET), Decoder(**VC2_VAE_DDCONFIG)
g = torch.Generator(device=dev).manual_seed(0)
for m in (unet, dec):
    m.to_empty(device=dev)
    with torch.no_grad():
        for name, p in m.named_parameters():
            if p.dim() >= 2:
                p.copy_(torch.randn(p.shape, generator=g, device=dev) * (0.6 / p[0].numel() ** 0.5))
            elif name.endswith("weight"):
                p.fill_(1.0)
            else:
                p.zero_()
    m.to(torch.bfloat16).eval()
unet.dtype = torch.bfloat16   # app.py:143
pq = torch.nn.Conv2d(4, 4, 1).to(dev, torch.bfloat16)


class FakeVAE(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.decoder, self.post_quant_conv = dec, pq

    def decode(self, z, **kw):   # autoencoder.py:110-113
        return self.decoder(self.post_quant_conv(z))


class FakeT2V(torch.nn.Module):   # the attributes the pipeline touches (pipeline:27-29,144,216)
    def __init__(self):
        super().__init__()
        self.first_stage_model = FakeVAE()
        self.model = torch.nn.Module()
        self.model.diffusion_model = unet
        self.cond_stage_model = torch.nn.Identity()
        self.temporal_length = FRAMES
        self.scale_factor = 0.18215

    def decode_first_stage_2DAE(self, z, **kw):   # ddpm3d.py:666-679
        z = 1.0 / self.scale_factor * z
        return torch.cat([self.first_stage_model.decode(z[:, :, i]).unsqueeze(2) for i in range(z.shape[2])], dim=2)


pipe = T2VTurboVC2Pipeline(FakeT2V(), T2VTurboScheduler(linear_start=0.00085, linear_end=0.012),
                           {"params": {"unet_config": {"params": VC2_UNET}}})
pe = torch.randn(a.batch, 77, 1024, device=dev, dtype=torch.bfloat16, generator=g)


def call():
    return pipe(prompt_embeds=pe, height=HEIGHT, width=WIDTH, frames=FRAMES, fps=16, guidance_scale=7.5,
                num_inference_steps=STEPS, lcm_origin_steps=50, generator=g, output_type="pt")


lat = torch.randn(a.batch, 4, FRAMES, HEIGHT // 8, WIDTH // 8, device=dev, dtype=torch.bfloat16, generator=g)
ts = torch.full((a.batch,), 999, device=dev, dtype=torch.long)
wemb = pipe.get_w_embedding(torch.tensor([7.5]).repeat(a.batch), 256).to(dev).to(torch.bfloat16)
for _ in range(2):
    vid = call()
assert tuple(vid.shape) == (a.batch, 3, FRAMES, HEIGHT, WIDTH)
torch.cuda.synchronize()
sampler = ClockSampler(0)
sampler.start()
times = []
for _ in range(a.calls):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    call()
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
ut = []
with torch.no_grad():
    for _ in range(a.calls):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        unet(lat, ts, context=pe, fps=16, timestep_cond=wemb)
        e1.record()
        torch.cuda.synchronize()
        ut.append(e0.elapsed_time(e1))
clocks = sampler.stop()
ms = statistics.median(times)
print(json.dumps(dict(what=f"unmodified reference pipeline (oracle/_ref), bf16, torch eager on B200, attention={a.attn}",
                      batch=a.batch, pipeline_4step_ms=ms, frames_per_s=FRAMES * a.batch / (ms * 1e-3), unet_fwd_ms=statistics.median(ut),
                      calls=a.calls, all_ms=[round(t, 2) for t in times], clocks=clocks, torch=torch.__version__)))
