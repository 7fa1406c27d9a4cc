"""One eager (no CUDA graph) 4-step pipeline call bracketed by cudaProfilerStart/Stop, for
`ncu --profile-from-start off` launch lists and single-kernel captures (see profiles/README.md)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import DEFAULT_BATCH, FRAMES, HEIGHT, WIDTH, STEPS, build_pipeline  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "pipeline"
BS = int(sys.argv[2]) if len(sys.argv) > 2 else DEFAULT_BATCH      # videos per call: the bench's throughput batch by default
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
pipe = build_pipeline(dev, use_graph=False)
gen = torch.Generator(device=dev).manual_seed(1)
pe = torch.randn(BS, 77, 1024, device=dev, dtype=torch.bfloat16, generator=gen)


def call():
    if what == "unet":
        lat = torch.randn(BS, 4, FRAMES, HEIGHT // 8, WIDTH // 8, device=dev, dtype=torch.bfloat16, generator=gen)
        ts = torch.full((BS,), 999, device=dev, dtype=torch.long)
        wemb = pipe.get_w_embedding(torch.tensor([7.5]).repeat(BS), 256).to(dev).to(torch.bfloat16)
        return pipe.unet(lat, ts, context=pe, fps=16, timestep_cond=wemb)
    return pipe(prompt_embeds=pe, height=HEIGHT, width=WIDTH, frames=FRAMES, fps=16, guidance_scale=7.5,
                num_inference_steps=STEPS, lcm_origin_steps=50, generator=gen, output_type="pt")


call()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
call()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one", what, "bs", BS)
