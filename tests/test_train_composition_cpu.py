"""The hand-written backward passes, checked WITHOUT a GPU: tests/mock_ops.py restates the contract of every kernel wrapper in
plain torch (fp32), it is patched over `t2v_turbo_b200.ops`, and the training views (StudentUNet: v1 LoRA step; FullUNet / V2Step:
v2 full fine-tune step) then run their real host composition on CPU.  Compared with

  * the UNMODIFIED reference's gradients (tests/golden/student_grads_small.pt, v2_step_small_motion.pt: oracle/make_goldens.py), and
  * autograd through the functional oracle (oracle/unet_oracle.py) for EVERY parameter,

at fp32 tolerances (1e-4): a wrong tape, layout, coefficient or a gradient routed to the wrong tensor cannot hide behind bf16
noise.  What this does not cover — that each CUDA kernel meets the contract restated in mock_ops — is the `-m gpu` suite's job.
"""
import os

import pytest
import torch

import mock_ops

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _rel(got, ref):
    got, ref = got.double(), ref.double()
    return ((got - ref).norm() / (ref.norm() + 1e-30)).item()


def _unet(name):
    from oracle.configs import UNET_CONFIGS
    from oracle.weights import seeded_state_dict
    from t2v_turbo_b200.unet import UNetModel
    spec = UNET_CONFIGS[name]
    m = UNetModel(**spec["cfg"])
    sd = seeded_state_dict(m.state_dict(), spec["weight_seed"])
    m.load_state_dict(sd, strict=True)
    return spec, m.eval(), sd


def test_student_unet_composition_vs_reference_lora_gradients(monkeypatch):
    """v1: StudentUNet forward + backward on CPU through the contract restatements == the reference's own LoRA gradients."""
    mock_ops.install(monkeypatch)
    from oracle.configs import student_loras, unet_inputs
    from t2v_turbo_b200.train_unet import StudentUNet
    g = torch.load(os.path.join(GOLD, "student_grads_small.pt"))
    spec, m, _ = _unet("small")
    s = StudentUNet(m, r=64, dropout_p=0.1, scale=1.0).eval()
    assert [tuple(x) for x in g["shapes"]] == s.arena.shapes
    s.arena.load_list(student_loras(g["shapes"]))
    s.pack()
    inp = unet_inputs(spec, g["timestep"])
    y = s(inp["x"], inp["timesteps"], context=inp["context"], fps=16, timestep_cond=inp["timestep_cond"])
    assert _rel(y, g["output"]) < 1e-4
    s.arena.zero_grad()
    s.backward(g["d_out"])
    n = len(s.arena.shapes)
    ratio = torch.tensor([s.arena.grad(i).double().norm().item() / max(g["grad_norms"][i].item(), 1e-30) for i in range(n)])
    assert (ratio - 1).abs().max() < 1e-4, (ratio.min(), ratio.max())
    worst = max(_rel(s.arena.grad(j), sc * t.float()) for j, (sc, t) in g["grads_full"].items())
    assert worst < 1e-3, worst          # the fixture keeps these tensors as fp16 scaled by their max


def test_student_unet_vc2_topology_vs_reference_lora_gradients(monkeypatch):
    """The same on the FULL VC2 topology (all 575 LoRA layers incl. the rank-4 conv_in / out, every level's stride-2 and upsampling
    convs, head counts 2/4/8/8) at 128 base channels: all 1150 LoRA gradient norms and 20 layers in full against the unmodified
    reference's autograd (tests/golden/student_grads_mid.pt)."""
    mock_ops.install(monkeypatch)
    from oracle.configs import student_loras, unet_inputs
    from t2v_turbo_b200.train_unet import StudentUNet
    g = torch.load(os.path.join(GOLD, "student_grads_mid.pt"))
    spec, m, _ = _unet("mid")
    spec = {**spec, "x_shape": tuple(g["x_shape"])}
    s = StudentUNet(m, r=64, dropout_p=0.1, scale=1.0).eval()
    assert [tuple(x) for x in g["shapes"]] == s.arena.shapes and len(s.arena.shapes) == 2 * 575
    s.arena.load_list(student_loras(g["shapes"]))
    s.pack()
    inp = unet_inputs(spec, g["timestep"])
    y = s(inp["x"], inp["timesteps"], context=inp["context"], fps=16, timestep_cond=inp["timestep_cond"])
    assert _rel(y, g["output"]) < 1e-4
    s.arena.zero_grad()
    s.backward(g["d_out"])
    n = len(s.arena.shapes)
    ratio = torch.tensor([s.arena.grad(i).double().norm().item() / max(g["grad_norms"][i].item(), 1e-30) for i in range(n)])
    assert (ratio - 1).abs().max() < 2e-4, (ratio.min(), ratio.max())
    worst = max(_rel(s.arena.grad(j), sc * t.float()) for j, (sc, t) in g["grads_full"].items())
    assert worst < 1e-3, worst


def test_param_groups_match_the_reference_rule():
    """full_train.param_groups == train_latent_t2v_turbo_v2.py:799-815 executed on the reference's own module tree (fixture)."""
    from t2v_turbo_b200.full_train import param_groups
    g = torch.load(os.path.join(GOLD, "v2_step_small_motion.pt"))
    _, m, _ = _unet("small_motion")
    assert [n for n, _ in m.named_parameters()] == g["names"]
    other, temporal = param_groups(m)
    assert temporal == g["temporal_names"]
    assert sorted(other + temporal) == sorted(g["names"]) and not set(other) & set(temporal)
    # the reference's quirk: the temporal transformer INSIDE middle_block is two components deep and lands in the other group
    assert any(n.startswith("init_attn.0") for n in temporal) and not any(n.startswith("middle_block") for n in temporal)


def test_full_unet_every_parameter_gradient_vs_oracle_autograd(monkeypatch):
    """v2: FullUNet (all 629 parameter tensors of the small motion-conditioned UNet train) vs autograd through the oracle."""
    mock_ops.install(monkeypatch)
    from oracle.configs import unet_inputs
    from oracle.unet_oracle import unet_forward
    from t2v_turbo_b200.full_train import FullUNet
    spec, m, sd = _unet("small_motion")
    s = FullUNet(m).eval()
    s.pack()
    inp = unet_inputs(spec, spec["timesteps"][0])
    kw = dict(fps=16, timestep_cond=inp["timestep_cond"], motion_cond=inp["motion_cond"])
    y = s(inp["x"], inp["timesteps"], context=inp["context"], **kw)
    sdg = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    yr = unet_forward(sdg, spec["cfg"], inp["x"], inp["timesteps"], inp["context"], **kw)
    assert _rel(y, yr.detach()) < 1e-4
    d_out = torch.randn(y.shape, generator=torch.Generator().manual_seed(4243))
    (yr * d_out).sum().backward()
    s.arena.zero_grad()
    s.backward(d_out)
    assert set(s.arena.names) == {k for k, v in sdg.items() if v.requires_grad}
    rels = {n: _rel(s.arena.grad(n), sdg[n].grad) for n in s.arena.names}
    worst = max(rels, key=rels.get)
    assert rels[worst] < 1e-4, (worst, rels[worst])
    # the nn.Parameters ARE the arena: state_dict() is the wire format, and a second backward accumulates
    assert all(p.data_ptr() == s.arena.param(n).data_ptr() for n, p in m.named_parameters())
    y2 = s(inp["x"], inp["timesteps"], context=inp["context"], **kw)
    s.backward(d_out)
    assert _rel(s.arena.grad(worst), 2 * sdg[worst].grad) < 1e-4 and torch.equal(y2, y)


def test_full_unet_vc2_topology_every_gradient_vs_oracle_autograd(monkeypatch):
    """The same on the FULL VC2 topology (all four levels, 12 input / 12 output blocks, head counts 2/4/8/8, stride-2 and upsampling
    convs at every level) at 128 base channels: all 1487 parameter tensors — the same tensor list and the same 33 optimizer-group runs
    as the 1.41 B model — against autograd through the oracle, batch 2 with per-sample timesteps and motion scales."""
    mock_ops.install(monkeypatch)
    from oracle.configs import UNET_CONFIGS, unet_inputs
    from oracle.unet_oracle import unet_forward
    from oracle.weights import seeded_state_dict
    from t2v_turbo_b200.full_train import FullUNet
    from t2v_turbo_b200.unet import UNetModel
    spec = dict(UNET_CONFIGS["mid"])
    spec.update(cfg={**spec["cfg"], "motion_cond_proj_dim": 256}, motion_gs=(0.05, 0.0), x_shape=(2, 4, 8, 16, 16))
    m = UNetModel(**spec["cfg"])
    sd = seeded_state_dict(m.state_dict(), 8)
    m.load_state_dict(sd, strict=True)
    s = FullUNet(m.eval()).eval()
    s.pack()
    assert len(s.arena.names) == 1487 and len(s.arena.runs) == 33
    inp = unet_inputs(spec, (759, 279))
    kw = dict(fps=16, timestep_cond=inp["timestep_cond"], motion_cond=inp["motion_cond"])
    y = s(inp["x"], inp["timesteps"], context=inp["context"], **kw)
    sdg = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    yr = unet_forward(sdg, spec["cfg"], inp["x"], inp["timesteps"], inp["context"], **kw)
    assert _rel(y, yr.detach()) < 1e-4
    d_out = torch.randn(y.shape, generator=torch.Generator().manual_seed(4243))
    (yr * d_out).sum().backward()
    s.arena.zero_grad()
    s.backward(d_out)
    rels = {n: _rel(s.arena.grad(n), sdg[n].grad) for n in s.arena.names}
    worst = max(rels, key=rels.get)
    assert rels[worst] < 1e-4, (worst, rels[worst])


def test_full_unet_gradients_vs_the_reference_itself(monkeypatch):
    """The same, against the UNMODIFIED reference's autograd (tests/golden/full_grads_small_motion.pt: linear loss, all 629 norms and
    57 tensors in full) — so the oracle used above is itself pinned on the reference for the backward, not only the forward."""
    mock_ops.install(monkeypatch)
    from oracle.configs import unet_inputs
    from t2v_turbo_b200.full_train import FullUNet
    g = torch.load(os.path.join(GOLD, "full_grads_small_motion.pt"))
    spec, m, _ = _unet("small_motion")
    assert [n for n, _ in m.named_parameters()] == g["names"]
    s = FullUNet(m).eval()
    s.pack()
    inp = unet_inputs(spec, g["timestep"])
    y = s(inp["x"], inp["timesteps"], context=inp["context"], fps=16, timestep_cond=inp["timestep_cond"], motion_cond=inp["motion_cond"])
    assert _rel(y, g["output"]) < 1e-4
    s.arena.zero_grad()
    s.backward(g["d_out"])
    ratio = torch.tensor([s.arena.grad(n).double().norm().item() / max(g["grad_norms"][n], 1e-30) for n in g["names"]])
    assert (ratio - 1).abs().max() < 1e-4, (ratio.min(), ratio.max())
    worst = max(_rel(s.arena.grad(n), sc * t.float()) for n, (sc, t) in g["grads_full"].items())
    assert worst < 1e-3, worst          # fp16-scaled storage


def test_full_unet_gradient_final_hooks_are_truthful(monkeypatch):
    """The data-parallel exchange trusts `on_grads_final(offset)`: every gradient at or above `offset` must already hold its final
    value when the hook fires (dist.ArenaReducer all-reduces those buckets while the backward continues), the offsets must fall
    monotonically to 0, and both optimizer groups must tile the arena as contiguous runs."""
    mock_ops.install(monkeypatch)
    from oracle.configs import unet_inputs
    from t2v_turbo_b200.full_train import FullUNet
    spec, m, _ = _unet("small_motion")
    s = FullUNet(m).eval()
    s.pack()
    a = s.arena
    assert a.runs[0][0] == 0 and a.runs[-1][1] == a.padded and all(x[1] == y[0] and x[2] != y[2] for x, y in zip(a.runs, a.runs[1:]))
    assert a.names[-1].startswith("out.") and a.names[0].startswith("time_embed")
    inp = unet_inputs(spec, spec["timesteps"][0])
    y = s(inp["x"], inp["timesteps"], context=inp["context"], fps=16, timestep_cond=inp["timestep_cond"], motion_cond=inp["motion_cond"])
    snaps = []
    s.on_grads_final = lambda off: snaps.append((off, a.grads[off:].clone()))
    a.zero_grad()
    s.backward(torch.randn(y.shape, generator=torch.Generator().manual_seed(1)))
    offs = [o for o, _ in snaps]
    assert offs == sorted(offs, reverse=True) and offs[-1] == 0 and len(set(offs)) >= 3 and offs[0] > a.padded // 2, offs
    for off, snap in snaps:
        assert torch.equal(a.grads[off:], snap), f"a gradient at or above offset {off} changed after it was reported final"


class _EvalTarget:
    """Stand-in for the EMA `UNetModel` (whose fused inference forward needs the GPU): the same weights through an eval-mode FullUNet."""
    dtype = torch.float32

    def __init__(self, view):
        self.view = view

    def __call__(self, *a, **k):
        y = self.view(*a, **k)
        self.view.detach_tapes()
        return y

    def invalidate_packed(self):
        pass


def _v2_setup(g, with_ema_target):
    from t2v_turbo_b200.distill_v2 import V2Step
    from t2v_turbo_b200.full_train import FullUNet
    from t2v_turbo_b200.scheduler import T2VTurboScheduler
    from t2v_turbo_b200.unet import UNetModel
    spec, m, sd = _unet("small_motion")
    s = FullUNet(m, with_target=True).eval()
    s.pack()
    target = None
    if with_ema_target:      # the fixture's EMA network: the student's weights times (1 + 0.02 N(0, 1)), drawn in parameter order
        gq = torch.Generator().manual_seed(g["target_perturb_seed"])
        m_t = UNetModel(**spec["cfg"])
        tsd = {}
        for n, p in m.named_parameters():
            tsd[n] = sd[n] * (1.0 + 0.02 * torch.randn(p.shape, generator=gq))
            s.arena.view(s.arena.target, s.arena.index[n]).copy_(tsd[n])
        m_t.load_state_dict({**sd, **tsd}, strict=True)
        tv = FullUNet(m_t.eval()).eval()
        tv.pack()
        target = _EvalTarget(tv)
    h = g["hyper"]
    step = V2Step(s, T2VTurboScheduler(linear_start=0.00085, linear_end=0.012), target_unet=target, num_ddim_timesteps=h["n_ddim"],
                  topk=h["topk"], motion_gs=h["motion_gs"], percentage=h["percentage"], use_motion_cond=True, loss_type="huber",
                  huber_c=0.001, timestep_scaling_factor=h["ts_scale"])
    return s, step, sd


def test_v2_step_vs_reference_composition(monkeypatch):
    """One v2 step (V2Step + train_step_v2: student forward with motion_cond, CFG + motion-prior guidance from the stored teacher
    outputs, DDIM step, EMA target, pseudo-Huber loss, full backward, clip, two-group AdamW, EMA update) == the same step composed
    from the UNMODIFIED reference's pieces (gen_v2_step)."""
    mock_ops.install(monkeypatch)
    from t2v_turbo_b200.distill_v2 import train_step_v2
    g = torch.load(os.path.join(GOLD, "v2_step_small_motion.pt"))
    s, step, sd = _v2_setup(g, with_ema_target=True)
    inp, h = g["inputs"], g["hyper"]
    batch = {k: inp[k] for k in ("index", "z_t", "cond_teacher_out", "uncond_teacher_out", "score", "use_motion_guide", "prompt_emb")}
    out = train_step_v2(step, batch, lr=h["lr"], temporal_lr_scale=h["temporal_lr_scale"], ema_decay=h["ema_decay"],
                        max_grad_norm=h["max_grad_norm"], weight_decay=h["weight_decay"], fixed=dict(w=inp["w"]))
    assert out["start_timesteps"].tolist() == g["start_timesteps"].tolist() and out["timesteps"].tolist() == g["timesteps"].tolist()
    assert torch.allclose(out["motion_gs"].float(), g["motion_gs"]) and g["motion_gs"].tolist()[1] == 0.0 and g["motion_gs"].tolist()[0] > 0
    for k in ("model_pred", "x_prev", "target"):
        assert _rel(out[k], g[k]) < 1e-4, (k, _rel(out[k], g[k]))
    assert abs(float(out["loss"]) - float(g["loss"])) < 1e-5 * float(g["loss"])
    # gradients: every norm, the stored tensors in full, the clip norm
    ratio = torch.tensor([s.arena.grad(n).double().norm().item() / max(g["grad_norms"][n], 1e-30) for n in s.arena.names])
    assert (ratio - 1).abs().max() < 2e-4, (ratio.min(), ratio.max())
    worst = max(_rel(s.arena.grad(n), sc * t.float()) for n, (sc, t) in g["grads_full"].items())
    assert worst < 1e-3, worst
    assert abs(float(s.arena.grad_norm()) - g["total_norm"]) < 1e-4 * g["total_norm"]
    # the optimizer step: parameter deltas of both lr groups (AdamW's first step moves every weight by ~lr: the temporal group 3x)
    temporal = set(g["temporal_names"])
    seen = set()
    for n, (sc, t) in g["param_delta"].items():
        d = s.arena.param(n) - sd[n]
        assert _rel(d, sc * t.float()) < 2e-3, (n, _rel(d, sc * t.float()))
        seen.add(n in temporal)
    assert seen == {True, False}, "the fixture must exercise both optimizer groups"
    # the EMA update of the target parameters
    for n, t in g["ema_after"].items():
        assert _rel(s.arena.view(s.arena.target, s.arena.index[n]), t) < 1e-6, n


def test_v2_step_self_target(monkeypatch):
    """--use_target_unet off: the student itself (gradient-free) is the target network (:1240)."""
    mock_ops.install(monkeypatch)
    g = torch.load(os.path.join(GOLD, "v2_step_small_motion.pt"))
    s, step, _ = _v2_setup(g, with_ema_target=False)
    inp = g["inputs"]
    batch = {k: inp[k] for k in ("index", "z_t", "cond_teacher_out", "uncond_teacher_out", "score", "prompt_emb")}   # use_motion_guide defaults to True
    s.arena.zero_grad()
    out = step(batch, fixed=dict(w=inp["w"]))
    assert _rel(out["target"], g["target_self"]) < 1e-4
    assert abs(float(out["loss"]) - float(g["loss_self_target"])) < 1e-5 * float(g["loss_self_target"])
    assert float(s.arena.grad_norm()) > 0


def test_v2_resume_reproduces_the_next_step(monkeypatch):
    """Save after step 1 (unet.state_dict() = the unet.pt wire format + FullArena.state_dict()), rebuild everything from the files,
    take step 2 in both: identical parameters, moments and EMA target."""
    mock_ops.install(monkeypatch)
    import io
    from t2v_turbo_b200.distill_v2 import train_step_v2
    g = torch.load(os.path.join(GOLD, "v2_step_small_motion.pt"))
    inp = g["inputs"]
    batch = {k: inp[k] for k in ("index", "z_t", "cond_teacher_out", "uncond_teacher_out", "score", "use_motion_guide", "prompt_emb")}
    kw = dict(lr=1e-4, temporal_lr_scale=2.0, ema_decay=0.9, fixed=dict(w=inp["w"]))
    s1, step1, _ = _v2_setup(g, with_ema_target=False)
    train_step_v2(step1, batch, **kw)
    buf = io.BytesIO()
    torch.save({"unet": s1.unet.state_dict(), "opt": s1.arena.state_dict()}, buf)
    train_step_v2(step1, batch, **kw)
    buf.seek(0)
    ck = torch.load(buf)
    s2, step2, _ = _v2_setup(g, with_ema_target=False)
    s2.unet.load_state_dict(ck["unet"], strict=True)          # writes straight into the arena (the nn.Parameters are its views)
    s2.arena.load_state_dict(ck["opt"])
    s2.refresh()
    train_step_v2(step2, batch, **kw)
    assert s2.arena.step == s1.arena.step == 2
    for a_, b_ in ((s1.arena.params, s2.arena.params), (s1.arena.exp_avg, s2.arena.exp_avg), (s1.arena.exp_avg_sq, s2.arena.exp_avg_sq),
                   (s1.arena.target, s2.arena.target)):
        assert torch.equal(a_, b_)


def test_v2_gradient_accumulation_and_video_reward_form(monkeypatch):
    """accumulate=True micro-batches add into the arena without stepping; the last call steps once with grad_scale = 1 / N: identical to
    one step on the mean gradient.  And the video-reward call form of reward_gradient (as_video: [B, F, 3, H, W], :1090-1094)."""
    mock_ops.install(monkeypatch)
    from t2v_turbo_b200.distill_v2 import train_step_v2
    g = torch.load(os.path.join(GOLD, "v2_step_small_motion.pt"))
    inp = g["inputs"]
    keys = ("index", "z_t", "cond_teacher_out", "uncond_teacher_out", "score", "use_motion_guide", "prompt_emb")
    halves = [{k: inp[k][i:i + 1] for k in keys} for i in (0, 1)]
    ws = [inp["w"][i:i + 1] for i in (0, 1)]
    s1, step1, _ = _v2_setup(g, with_ema_target=False)
    p0 = s1.arena.params.clone()
    train_step_v2(step1, halves[0], lr=1e-4, accumulate=True, max_grad_norm=None, fixed=dict(w=ws[0]))
    assert torch.equal(s1.arena.params, p0) and s1.arena.step == 0 and float(s1.arena.grad_norm()) > 0
    g_first = s1.arena.grads.clone()
    train_step_v2(step1, halves[1], lr=1e-4, grad_scale=0.5, max_grad_norm=None, fixed=dict(w=ws[1]))
    assert s1.arena.step == 1 and not torch.equal(s1.arena.grads, g_first)
    # reference: the two micro-batch gradients summed by hand, one AdamW step on their mean
    s2, step2, _ = _v2_setup(g, with_ema_target=False)
    s2.arena.zero_grad()
    step2(halves[0], fixed=dict(w=ws[0]))
    step2(halves[1], fixed=dict(w=ws[1]))
    assert _rel(s1.arena.grads, s2.arena.grads) < 1e-6
    s2.arena.adamw_step(lr=1e-4, weight_decay=0.0, grad_scale=0.5)
    assert torch.allclose(s1.arena.params, s2.arena.params, rtol=0, atol=1e-9)
    train_step_v2(step1, halves[0], lr=1e-4, accumulate=True, max_grad_norm=None, fixed=dict(w=ws[0]))   # a new accumulation starts from zero
    g_new = s1.arena.grads.clone()
    s1.arena.zero_grad()
    step1(halves[0], fixed=dict(w=ws[0]))
    assert torch.equal(g_new, s1.arena.grads)
    # video-reward form
    from oracle.configs import VAE_CONFIGS
    from oracle.weights import vae_state_dict
    from t2v_turbo_b200.vae import AutoencoderKL
    from t2v_turbo_b200.vae_train import reward_gradient
    vspec = VAE_CONFIGS["small"]
    vae = AutoencoderKL(vspec["ddconfig"], vspec["embed_dim"])
    vae.load_state_dict(vae_state_dict(vae.state_dict(), vspec["weight_seed"]))
    seen = {}

    def video_rm(imgs):
        seen["shape"] = tuple(imgs.shape)
        return imgs.mean((1, 2, 3, 4))
    mp_ = torch.randn(2, 4, 4, 8, 8, generator=torch.Generator().manual_seed(9))
    loss, d = reward_gradient(vae.eval(), mp_, video_rm, frame_idx=[0, 2, 3], batch_idx=[1], reward_scale=2.0, as_video=True)
    assert seen["shape"] == (1, 3, 3, 32, 32) and d.shape == mp_.shape
    assert d[0].abs().max() == 0 and d[1][:, 1].abs().max() == 0 and d[1][:, 0].abs().max() > 0 and torch.isfinite(loss)


def test_attach_ema_target_aliases_the_target_arena(monkeypatch):
    """The EMA network's nn.Parameters are views of arena.target: an `ema_step` changes what its next forward packs, its state_dict()
    is the EMA checkpoint, and nothing aliases the student's live parameters."""
    mock_ops.install(monkeypatch)
    from t2v_turbo_b200.distill_v2 import attach_ema_target
    from t2v_turbo_b200.full_train import FullUNet
    from t2v_turbo_b200.unet import UNetModel
    spec, m, sd = _unet("small_motion")
    s = FullUNet(m, with_target=True)
    t = attach_ema_target(s, UNetModel(**spec["cfg"]))
    assert t.dtype == torch.bfloat16 and not t.training
    a = s.arena
    for n, p in t.named_parameters():
        assert p.data_ptr() == a.view(a.target, a.index[n]).data_ptr() and torch.equal(p, sd[n])
    a.params.add_(1.0)
    a.ema_step(0.75)
    n0 = a.names[5]
    assert torch.allclose(dict(t.named_parameters())[n0], sd[n0] + 0.25) and torch.allclose(t.state_dict()[n0], sd[n0] + 0.25)
    with pytest.raises(RuntimeError):
        attach_ema_target(FullUNet(_unet("small_motion")[1]), UNetModel(**spec["cfg"]))


def test_v2_host_draws_motion_condition():
    """The motion-guidance gate and coefficient on the host (:1019-1031, :1214-1226), incl. the reference's sqrt(1 - sqrt(alpha_bar))."""
    from t2v_turbo_b200.distill_v2 import V2Step
    from t2v_turbo_b200.scheduler import T2VTurboScheduler
    sch = T2VTurboScheduler(linear_start=0.00085, linear_end=0.012)
    step = V2Step(None, sch, num_ddim_timesteps=200, topk=5, motion_gs=0.05, percentage=0.5)
    idx = torch.tensor([199, 100, 99, 150])
    H = step.host_draws(idx, torch.tensor([True, True, True, False]), fixed=dict(w=torch.tensor([5.0, 6.0, 7.0, 8.0])))
    assert H["motion_gs"].tolist() == [0.05, 0.05, 0.0, 0.0]                     # index >= 100 and use_motion_guide
    ac = sch.alphas_cumprod.double()
    t = H["start_timesteps"]
    assert t.tolist() == [999, 504, 499, 754] and H["timesteps"].tolist() == [994, 499, 494, 749]
    exp = -0.05 * (1 - ac[t].sqrt()).sqrt()
    assert torch.allclose(H["mg"][:2].double(), exp[:2], rtol=1e-6) and H["mg"][2:].abs().max() == 0
    assert H["mg_emb"].shape == (4, 256) and torch.equal(H["mg_emb"][2], H["mg_emb"][3])
    step_nc = V2Step(None, sch, use_motion_cond=False)
    assert "mg_emb" not in step_nc.host_draws(idx, torch.ones(4, dtype=torch.bool))


# ----------------------------------------------------------------------------- vae.decode WITH grad (the reward terms' path)
def test_decoder_grad_vs_oracle_autograd(monkeypatch):
    """vae_train.DecoderGrad: forward == the VAE oracle, the hand-written input gradient == autograd through it (incl. the AttnBlock's
    softmax adjoint, the upsampling adjoints, the padded 3- / 4-channel ends, post_quant_conv and the latent scale), through the
    public `decode_with_grad` in both call forms, with a non-linear "reward" on top."""
    mock_ops.install(monkeypatch)
    from oracle.configs import VAE_CONFIGS
    from oracle.vae_oracle import decode_first_stage_2dae
    from oracle.weights import vae_state_dict
    from t2v_turbo_b200.vae import AutoencoderKL
    from t2v_turbo_b200.vae_train import decode_with_grad
    spec = VAE_CONFIGS["small"]
    v = AutoencoderKL(spec["ddconfig"], spec["embed_dim"])
    sd = vae_state_dict(v.state_dict(), spec["weight_seed"])
    v.load_state_dict(sd)
    v.eval()
    z = torch.randn(spec["z_shape"], generator=torch.Generator().manual_seed(3))                # [1, 4, 4, 16, 16]
    side = 16 * 2 ** (len(spec["ddconfig"]["ch_mult"]) - 1)
    probe = torch.randn(1, 3, 4, side, side, generator=torch.Generator().manual_seed(5))

    def reward(img):            # stands in for a reward model: clamp to [0, 1] like the scripts, then a non-linear score
        x = (img / 2 + 0.5).clamp(0, 1)
        return (x * probe).sum() + (x ** 2).mean()
    z1 = z.clone().requires_grad_(True)
    img = decode_with_grad(v, z1, scale=1.0 / 0.18215)
    reward(img).backward()
    z2 = z.clone().requires_grad_(True)
    ref = decode_first_stage_2dae(sd, spec["ddconfig"], z2)
    reward(ref).backward()
    assert _rel(img.detach(), ref.detach()) < 1e-4, _rel(img.detach(), ref.detach())
    assert _rel(z1.grad, z2.grad) < 1e-4, _rel(z1.grad, z2.grad)
    # the training scripts' call form: frames [N, zc, h, w] (selected_latents, train_t2v_turbo_v1_lora.py:1055-1062)
    zf = z[0].permute(1, 0, 2, 3).contiguous().requires_grad_(True)
    imgf = decode_with_grad(v, zf, scale=1.0 / 0.18215)
    assert imgf.shape == (4, 3, side, side) and _rel(imgf.detach(), ref.detach()[0].permute(1, 0, 2, 3)) < 1e-4
    imgf.square().sum().backward()
    assert zf.grad.shape == zf.shape and torch.isfinite(zf.grad).all() and zf.grad.abs().max() > 0


def test_v2_step_with_reward_branch_vs_oracle_autograd(monkeypatch):
    """The step with a reward term (train_latent_t2v_turbo_v2.py:1062-1098 shape: frames of model_pred -> vae.decode -> clamp -> reward):
    `V2Step.reward = partial(vae_train.reward_gradient, vae, reward_fn=...)` adds the reward's gradient — through the hand-written decoder
    adjoint — to the distillation gradient before the student backward.  Every parameter gradient vs autograd of the SAME total loss
    composed from the UNet oracle and the VAE oracle."""
    import functools
    mock_ops.install(monkeypatch)
    from oracle.configs import UNET_CONFIGS, VAE_CONFIGS
    from oracle.unet_oracle import unet_forward
    from oracle.vae_oracle import decode_first_stage_2dae
    from oracle.weights import vae_state_dict
    from t2v_turbo_b200.vae import AutoencoderKL
    from t2v_turbo_b200.vae_train import reward_gradient
    g = torch.load(os.path.join(GOLD, "v2_step_small_motion.pt"))
    s, step, sd = _v2_setup(g, with_ema_target=False)
    vspec = VAE_CONFIGS["small"]
    vae = AutoencoderKL(vspec["ddconfig"], vspec["embed_dim"])
    vsd = vae_state_dict(vae.state_dict(), vspec["weight_seed"])
    vae.load_state_dict(vsd)
    vae.eval()
    frame_idx, scale = [2, 0], 0.37

    def reward_fn(imgs):                      # any differentiable score of images in [0, 1]
        return -((imgs - 0.3) ** 2).mean((1, 2, 3))
    step.reward = functools.partial(reward_gradient, vae, reward_fn=reward_fn, frame_idx=frame_idx, reward_scale=scale)
    inp = g["inputs"]
    batch = {k: inp[k] for k in ("index", "z_t", "cond_teacher_out", "uncond_teacher_out", "score", "use_motion_guide", "prompt_emb")}
    s.arena.zero_grad()
    out = step(batch, fixed=dict(w=inp["w"]))
    # the same total loss under autograd
    spec = UNET_CONFIGS["small_motion"]
    H = step.host_draws(inp["index"], inp["use_motion_guide"], fixed=dict(w=inp["w"]))
    sdg = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    eps = unet_forward(sdg, spec["cfg"], inp["z_t"], H["start_timesteps"], inp["prompt_emb"], fps=16, timestep_cond=H["w_emb"], motion_cond=H["mg_emb"])
    v5 = lambda t: t.view(-1, 1, 1, 1, 1)      # noqa: E731
    model_pred = v5(H["k_z"]) * inp["z_t"] + v5(H["k_e"]) * eps
    d = model_pred - out["target"]
    distill = torch.mean(torch.sqrt(d ** 2 + 0.001 ** 2) - 0.001)
    img = decode_first_stage_2dae(vsd, vspec["ddconfig"], model_pred[:, :, frame_idx])
    imgs = (img / 2 + 0.5).clamp(0, 1).permute(0, 2, 1, 3, 4).reshape(-1, *img.shape[1:2], *img.shape[3:])
    r_loss = -reward_fn(imgs).mean() * scale
    (distill + r_loss).backward()
    assert _rel(out["model_pred"], model_pred.detach()) < 1e-4
    r_ref, d_ref = float(r_loss.detach()), float(distill.detach())
    assert abs(float(out["reward_loss"]) - r_ref) < 1e-5 * abs(r_ref) and abs(float(out["loss"]) - d_ref) < 1e-5 * d_ref
    rels = {n: _rel(s.arena.grad(n), sdg[n].grad) for n in s.arena.names}
    worst = max(rels, key=rels.get)
    assert rels[worst] < 2e-4, (worst, rels[worst])
    # and the reward really contributes: without it the gradient differs
    s2, step2, _ = _v2_setup(g, with_ema_target=False)
    s2.arena.zero_grad()
    step2(batch, fixed=dict(w=inp["w"]))
    assert _rel(s2.arena.grads, s.arena.grads) > 1e-3


def test_v1_distill_step_vs_reference_composition(monkeypatch):
    """The v1 step (distill.DistillStep: add_noise, LoRA student, batched teacher CFG + DDIM step, gradient-free target, pseudo-Huber loss,
    student backward) in fp32 on CPU == the step composed from the UNMODIFIED reference's pieces (tests/golden/distill_step_small.pt).  On
    the GPU this comparison is noise-dominated (sign-like loss gradient in bf16, tests/test_student_gpu.py); here it is exact to 1e-4."""
    mock_ops.install(monkeypatch)
    from oracle.configs import UNET_CONFIGS, student_loras
    from oracle.weights import seeded_state_dict
    from t2v_turbo_b200.distill import DistillStep
    from t2v_turbo_b200.scheduler import T2VTurboScheduler
    from t2v_turbo_b200.train_unet import StudentUNet
    from t2v_turbo_b200.unet import UNetModel
    g = torch.load(os.path.join(GOLD, "distill_step_small.pt"))
    spec = UNET_CONFIGS["small"]
    base = UNetModel(**spec["cfg"])
    sd = seeded_state_dict(base.state_dict(), spec["weight_seed"])
    base.load_state_dict(sd, strict=True)
    tcfg = dict(spec["cfg"])
    tcfg["time_cond_proj_dim"] = None
    teacher_m = UNetModel(**tcfg)
    teacher_m.load_state_dict({k: v for k, v in sd.items() if not k.startswith("time_cond_proj")}, strict=True)
    s = StudentUNet(base.eval(), r=64).eval()
    assert [tuple(x) for x in g["shapes"]] == s.arena.shapes
    s.arena.load_list(student_loras(g["shapes"]))
    s.pack()
    tv = StudentUNet(teacher_m.eval(), r=64).eval()          # lora_up = 0 at initialisation: exactly the frozen teacher's forward
    tv.pack()
    step = DistillStep(s, _EvalTarget(tv), T2VTurboScheduler(linear_start=0.00085, linear_end=0.012), num_ddim_timesteps=50, topk=20,
                       loss_type="huber", huber_c=0.001, timestep_scaling_factor=10.0)
    inp = g["inputs"]
    s.arena.zero_grad()
    out = step(inp["latents"], inp["prompt"], inp["uncond"], fixed=dict(index=inp["index"], noise=inp["noise"], w=inp["w"]))
    assert out["start_timesteps"].tolist() == g["start_timesteps"].tolist() and out["timesteps"].tolist() == g["timesteps"].tolist()
    for k in ("model_pred", "x_prev", "target"):
        assert _rel(out[k], g[k]) < 1e-4, (k, _rel(out[k], g[k]))
    assert abs(float(out["loss"]) - float(g["loss"])) < 1e-5 * float(g["loss"])
    n = len(s.arena.shapes)
    ratio = torch.tensor([s.arena.grad(i).double().norm().item() / max(g["grad_norms"][i].item(), 1e-30) for i in range(n)])
    assert (ratio - 1).abs().max() < 5e-4, (ratio.min(), ratio.max())
    worst = max(_rel(s.arena.grad(j), sc * t.float()) for j, (sc, t) in g["grads_full"].items())
    assert worst < 2e-3, worst


# ----------------------------------------------------------------------------- the motion-prior score (v2 preprocessing)
def test_motion_prior_score_vs_reference_autograd(monkeypatch):
    """motion_prior.get_motion_prior_score (ScoreUNet: probabilities exported in the forward, their gradient injected into the temporal
    attention adjoints, input gradient through conv_in) == the UNMODIFIED reference's `torch.autograd.grad(loss, latents)` with its own
    compute_temp_loss (tests/golden/motion_score_small.pt)."""
    mock_ops.install(monkeypatch)
    from oracle.weights import seeded_state_dict
    from oracle.configs import UNET_CONFIGS
    from t2v_turbo_b200.motion_prior import ScoreUNet, get_motion_prior_score, temp_loss_and_grad
    from t2v_turbo_b200.unet import UNetModel
    g = torch.load(os.path.join(GOLD, "motion_score_small.pt"))
    m = UNetModel(**g["cfg"])
    m.load_state_dict(seeded_state_dict(m.state_dict(), UNET_CONFIGS["small"]["weight_seed"]), strict=True)
    view = ScoreUNet(m.eval())
    assert view.probe_names == g["layers"]
    view.pack()
    score, eps = get_motion_prior_score(view, g["latents"], g["ts"], g["example"], {"context": g["ctx_orig"], "fps": 16},
                                        {"context": g["ctx_inf"], "fps": 16}, g["temp_loss_scale"])
    assert _rel(eps, g["cond_teacher_output"]) < 1e-4
    assert score.shape == g["score"].shape and _rel(score, g["score"]) < 2e-4, _rel(score, g["score"])
    # the exported probabilities and the loss value
    _, probs = view(g["latents"], g["ts"], context=g["ctx_inf"], fps=16)
    view.detach_tapes()
    for n in g["layers"]:
        assert (probs[n] - g["probs"][n].float()).abs().max() < 1e-3           # fixture keeps them in fp16
    _, pe = view(g["example"], g["ts"], context=g["ctx_orig"], fps=16)
    view.detach_tapes()
    loss, _ = temp_loss_and_grad(probs, pe, g["temp_loss_scale"])
    assert abs(float(loss) - float(g["loss"])) < 1e-4 * float(g["loss"])


def test_preprocess_sample_vs_reference_composition(monkeypatch):
    """motion_prior.preprocess_sample — add_noise, the DDIM inversion loop (index + 1 teacher forwards), the unconditional teacher output,
    the motion-prior score — == preprocess_with_motion_prior.py:326-401 composed from the UNMODIFIED reference's pieces, and the record
    survives the latent dataset's wire format (formats.dumps_v2_sample / loads_v2_sample, fp16)."""
    mock_ops.install(monkeypatch)
    from oracle.configs import UNET_CONFIGS
    from oracle.weights import seeded_state_dict
    from t2v_turbo_b200 import formats
    from t2v_turbo_b200.distill import DDIMSolver
    from t2v_turbo_b200.motion_prior import ScoreUNet, preprocess_sample
    from t2v_turbo_b200.scheduler import T2VTurboScheduler
    from t2v_turbo_b200.unet import UNetModel
    g = torch.load(os.path.join(GOLD, "preprocess_sample_small.pt"))
    m = UNetModel(**g["cfg"])
    m.load_state_dict(seeded_state_dict(m.state_dict(), UNET_CONFIGS["small"]["weight_seed"]), strict=True)
    view = ScoreUNet(m.eval())
    view.pack()
    sch = T2VTurboScheduler(linear_start=0.00085, linear_end=0.012)
    solver = DDIMSolver(sch.alphas_cumprod.numpy(), ddim_timesteps=50)
    ref = g["record"]

    class _AddNoise:          # T2VTurboScheduler.add_noise refuses CPU tensors (its kernel is GPU-only and GPU-tested): closed form here
        @staticmethod
        def add_noise(x, n, t):
            a = sch.alphas_cumprod[t.cpu()].view(-1, 1, 1, 1, 1)
            return a.sqrt() * x + (1 - a).sqrt() * n
    rec = preprocess_sample(view, _AddNoise, solver, g["latents"], g["prompt"], g["uncond"], index=ref["index"], noise=g["noise"],
                            temp_loss_scale=g["temp_loss_scale"])
    assert set(rec) == set(formats.V2_SAMPLE_KEYS) and int(rec["index"]) == int(ref["index"])
    for k in formats.V2_SAMPLE_KEYS:
        if k != "index":
            assert _rel(rec[k], ref[k]) < 3e-4, (k, _rel(rec[k], ref[k]))
    back = formats.loads_v2_sample(formats.dumps_v2_sample(**rec), frames=g["latents"].shape[2])
    assert _rel(back["score"].float(), ref["score"]) < 2e-3 and back["z_t"].dtype == torch.float16
