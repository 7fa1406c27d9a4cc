"""ModelScope path (BASELINE config 5).  PARITY UNPINNED: diffusers is absent, so the checker is
oracle/ms_oracle.py — a restatement of diffusers 0.30.0's blocks wired as model_scope/unet_3d_blocks.py does — plus an
independent parameter census of UNet3DConditionModel.  CPU tests: key maps / census; GPU tests: the B200 adapter vs the oracle."""
import math

import pytest
import torch

SMALL = dict(block_out_channels=(64, 128), layers_per_block=1, cross_attention_dim=128)


def _seeded(shapes, seed):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k in sorted(shapes):
        s = shapes[k]
        if len(s) >= 2:
            sd[k] = torch.randn(s, generator=g) * (0.8 / math.sqrt(math.prod(s[1:])))
        elif k.endswith("weight"):
            sd[k] = 1.0 + 0.1 * torch.randn(s, generator=g)
        else:
            sd[k] = 0.05 * torch.randn(s, generator=g)
    return sd


def test_ms_unet_key_map_matches_independent_census():
    """Every diffusers key of UNet3DConditionModel (census written from its constructors, oracle/ms_oracle.py) maps onto exactly
    one parameter of the B200 UNetModel with the same shape: 1481 tensors / 1 411 315 780 parameters at the full config."""
    from oracle.ms_oracle import ms_unet_param_shapes
    from t2v_turbo_b200.ms_adapter import UNet3DConditionModel, convert_ms_unet_state_dict
    for kw, n_params in ((dict(), 1_411_315_780), (SMALL, None)):
        shapes = ms_unet_param_shapes(**kw)
        with torch.device("meta"):
            m = UNet3DConditionModel(**kw)
            sd = {k: torch.empty(v) for k, v in shapes.items()}
        conv = convert_ms_unet_state_dict(sd, kw.get("block_out_channels", (320, 640, 1280, 1280)), kw.get("layers_per_block", 2))
        own = m.model.state_dict()
        assert set(conv) == set(own)
        for k, v in conv.items():
            assert v.numel() == own[k].numel() and (tuple(v.shape) == tuple(own[k].shape)), k
        if n_params:
            assert len(shapes) == 1481 and sum(math.prod(s) for s in shapes.values()) == n_params
    with pytest.raises(KeyError):
        convert_ms_unet_state_dict({"down_blocks.0.bogus.weight": torch.zeros(1)})


def test_diffusers_vae_key_map_covers_the_kl_vae():
    """diffusers AutoencoderKL names -> lvdm names: a bijection onto the B200 AutoencoderKL's parameters (up blocks are
    indexed by execution order in diffusers and by resolution level in lvdm; the mid attention is Linear vs 1x1 conv)."""
    from t2v_turbo_b200.configs import VC2_VAE_DDCONFIG
    from t2v_turbo_b200.ms_adapter import convert_diffusers_vae_state_dict, diffusers_vae_key_map
    from t2v_turbo_b200.vae import AutoencoderKL
    with torch.device("meta"):
        own = AutoencoderKL(VC2_VAE_DDCONFIG, 4).state_dict()
    inv = {v: k for k, v in diffusers_vae_key_map().items()}
    src = {}
    for k, v in own.items():
        pref = max((p for p in inv if k == p or k.startswith(p + ".")), key=len)
        w = torch.empty(v.shape[:2], device="meta") if (".mid.attn_1." in k and k.endswith("weight") and v.dim() == 4) else v
        src[inv[pref] + k[len(pref):]] = w
    assert "decoder.up_blocks.0.resnets.0.conv1.weight" in src and "decoder.mid_block.attentions.0.to_q.weight" in src
    conv = convert_diffusers_vae_state_dict(src)
    assert set(conv) == set(own) and all(tuple(conv[k].shape) == tuple(own[k].shape) for k in own)
    assert inv["decoder.up.3.block.0.conv1"] == "decoder.up_blocks.0.resnets.0.conv1"      # lowest resolution runs first
    assert inv["decoder.up.0.block.2.conv2"] == "decoder.up_blocks.3.resnets.2.conv2"


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_ms_unet_vs_oracle(cuda_device, dtype):
    """B200 UNet3DConditionModel (diffusers-keyed weights, reference call form) against the fp32 restatement; fp16 tensors at
    the boundary as app_ms.py:67-69 runs it."""
    from oracle.ms_oracle import ms_unet_param_shapes, unet3d_forward
    from t2v_turbo_b200.ms_adapter import UNet3DConditionModel
    sd = _seeded(ms_unet_param_shapes(**SMALL), 31)
    m = UNet3DConditionModel(**SMALL)
    m.load_state_dict(sd)                      # diffusers key names
    m = m.cuda().eval()
    g = torch.Generator().manual_seed(32)
    x = torch.randn(2, 4, 4, 8, 8, generator=g)
    ctx = torch.randn(2, 77, 128, generator=g)
    ts = torch.tensor([759, 279])
    w = torch.randn(2, 256, generator=g)
    with torch.no_grad():
        ref = unet3d_forward(sd, SMALL, x, ts, ctx, timestep_cond=w)
    out = m(x.to(dtype).cuda(), ts.cuda(), encoder_hidden_states=ctx.to(dtype).cuda(), timestep_cond=w.to(dtype).cuda()).sample
    assert out.dtype == dtype and out.shape == ref.shape
    rel = ((out.float().cpu() - ref).norm() / ref.norm()).item()
    print(f"\n[ms unet {dtype}] rel-L2 vs diffusers-restatement oracle {rel:.3e}")
    assert rel <= 3.0e-2


@pytest.mark.gpu
def test_ms_pipeline_runs_and_matches_oracle_composition(cuda_device):
    """T2VTurboMSPipeline call surface (t2v_turbo_ms_pipeline.py:132-221), 4 steps, small config: latents vs the loop composed
    from the oracles with the same noise (the MS pipeline draws step noise from the global RNG: seeded here)."""
    from oracle.configs import VAE_CONFIGS
    from oracle.ms_oracle import ms_unet_param_shapes, unet3d_forward
    from oracle.scheduler_oracle import SchedulerOracle
    from oracle.unet_oracle import guidance_scale_embedding
    from oracle.weights import vae_state_dict
    from t2v_turbo_b200.ms_adapter import DiffusersAutoencoderKL, T2VTurboMSPipeline, UNet3DConditionModel
    from t2v_turbo_b200.scheduler import T2VTurboScheduler
    sd = _seeded(ms_unet_param_shapes(**SMALL), 33)
    unet = UNet3DConditionModel(**SMALL)
    unet.load_state_dict(sd)
    vspec = VAE_CONFIGS["small"]
    vae = DiffusersAutoencoderKL(vspec["ddconfig"])
    vae.vae.load_state_dict(vae_state_dict(vae.vae.state_dict(), vspec["weight_seed"]))
    pe = torch.randn(1, 77, 128, generator=torch.Generator().manual_seed(34))
    lat0 = torch.randn(1, 4, 4, 8, 8, generator=torch.Generator().manual_seed(35))
    noises = [torch.randn(1, 4, 4, 8, 8, generator=torch.Generator().manual_seed(36 + i)) for i in range(4)]
    for graph in (False, True):
        pipe = T2VTurboMSPipeline(unet.cuda().eval(), vae.cuda().eval(), scheduler=T2VTurboScheduler(linear_start=0.00085, linear_end=0.012),
                                  use_cuda_graph=graph)
        # our scheduler accepts the noise explicitly (variance_noise); feed the same tensors through a tiny generator shim
        it = iter(noises)
        orig = pipe.scheduler.step
        pipe.scheduler.step = lambda *a, **k: orig(*a, **{**k, "variance_noise": next(it).cuda()})
        den = pipe(prompt_embeds=pe.cuda(), height=64, width=64, frames=4, num_inference_steps=4, latents=lat0.clone(), output_type="latent")
        it = iter(noises)
        vid = pipe(prompt_embeds=pe.cuda(), height=64, width=64, frames=4, num_inference_steps=4, latents=lat0.clone(), output_type="pt")
        assert vid.shape == (1, 3, 4, 32, 32)
        so = SchedulerOracle()
        tsx = so.set_timesteps(4, 50)
        lat, ref_den = lat0.clone(), None
        w = guidance_scale_embedding(torch.tensor([7.5]), 256)
        with torch.no_grad():
            for i, t in enumerate(tsx):
                pred = unet3d_forward(sd, SMALL, lat, torch.tensor([int(t)]), pe, timestep_cond=w)
                lat, ref_den = so.step(pred, i, int(t), lat, noise=noises[i])
        rel = ((den.float().cpu() - ref_den).norm() / ref_den.norm()).item()
        print(f"\n[ms pipeline graph={graph}] denoised latent rel-L2 vs oracle loop {rel:.3e}")
        assert rel <= 2.0e-2 and torch.isfinite(vid.float()).all()
