"""KL-VAE encode timing: 16 frames 320x512 -> latent [1,4,16,40,64]; CUDA events, eager."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from t2v_turbo_b200.configs import VC2_VAE_DDCONFIG
from t2v_turbo_b200.vae import AutoencoderKL
from t2v_turbo_b200 import ops
dev = torch.device("cuda", 0)
m = AutoencoderKL(VC2_VAE_DDCONFIG, 4).to(dev).eval()
x = torch.randn(1, 3, 16, 320, 512, device=dev, dtype=torch.bfloat16)
noise = torch.randn(16, 4, 40, 64)
for _ in range(3):
    z = m.encode_frames(x, noise=noise, scale=0.18215)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n0 = ops.LAUNCHES
e0.record()
for _ in range(5):
    z = m.encode_frames(x, noise=noise, scale=0.18215)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"encode 16x320x512: {ms:.2f} ms per video ({16 / ms * 1e3:.1f} frames/s), 11.04 TFLOP -> {11.04 / ms * 1e3:.0f} TFLOP/s, "
      f"{(ops.LAUNCHES - n0) // 5} launches, finite={bool(torch.isfinite(z.float()).all())}")
