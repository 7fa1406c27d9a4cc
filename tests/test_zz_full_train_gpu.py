"""GPU parity of the v2 full fine-tune step (t2v_turbo_b200/full_train.py, distill_v2.py, csrc/train_full.cu).

STATUS — read before trusting a green line here: this file was written after the round's GPU budget was spent and has NEVER
been executed on a B200.  Its host composition is verified on CPU (tests/test_train_composition_cpu.py: every parameter gradient
against autograd, the whole step against the unmodified reference's composition); what remains unproven is that the three new
kernels and t2v_wgrad's in-place 64-column B slices meet their contracts on the device.  Therefore every test is a NON-STRICT
xfail: it runs, a pass is reported as XPASS and a failure as xfail, and neither turns the suite red; the file sorts last so that
nothing runs after it in the same process.  The bounds are first guesses from the v1 student tests, not observed errors.

Kernel contracts are checked against tests/mock_ops.py evaluated on the CPU in fp32 on the same bf16-rounded inputs.
"""
import os

import pytest
import torch

import mock_ops
from test_kernels_gpu import BF16, _ops, assert_close, rnd

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason="v2 full fine-tune step: never run on a GPU (round-2 budget exhausted); CPU-verified composition")]
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _rel(got, ref):
    got, ref = got.double().cpu(), ref.double().cpu()
    return ((got - ref).norm() / (ref.norm() + 1e-30)).item()


def _mock(dtype=torch.float32):
    mock_ops.ACT = dtype
    return mock_ops


# ----------------------------------------------------------------------------- the new kernels
@pytest.mark.parametrize("n,hw,c,silu", [(4, 640, 320, True), (2, 2560, 640, True), (3, 160, 1280, False), (16, 40, 128, True),
                                          (2, 1000, 64, False), (1, 320, 2560, True)])
def test_groupnorm_affine_grad(cuda_device, n, hw, c, silu):
    ops = _ops()
    x = (rnd(n * hw, c, scale=1.5, seed=1) + 0.7).to(BF16)
    dy = rnd(n * hw, c, seed=2).to(BF16)
    gamma, beta = (1.0 + 0.3 * rnd(c, seed=3)).float(), (0.2 * rnd(c, seed=4)).float()
    dg = torch.full((c,), 0.5, device="cuda")          # accumulated INTO
    db = torch.full((c,), -0.25, device="cuda")
    ws = []
    ops.groupnorm_bwd(x, dy, gamma, beta, rows_per_sample=hw, eps=1e-5, silu=silu, keep_ws=ws)
    ops.groupnorm_affine_grad(x, dy, gamma, beta, ws[0], dg, db, rows_per_sample=hw, eps=1e-5, silu=silu)
    m = _mock()
    rg, rb = torch.full((c,), 0.5), torch.full((c,), -0.25)
    xs = x.float().cpu()
    m.groupnorm_affine_grad(xs, dy.float().cpu(), gamma.cpu(), beta.cpu(), ("gn-stats", xs.data_ptr(), hw, 32), rg, rb,
                            rows_per_sample=hw, eps=1e-5, silu=silu)
    assert_close(dg.cpu(), rg, rtol=2e-3, atol_scale=1e-3, what=f"gn dgamma n={n} hw={hw} c={c}")     # fp32 sums: tighter than bf16 outputs
    assert_close(db.cpu(), rb, rtol=2e-3, atol_scale=1e-3, what=f"gn dbeta n={n} hw={hw} c={c}")


@pytest.mark.parametrize("rows,c", [(4096, 320), (2560, 640), (1000, 1280), (77, 64), (5, 1024)])
def test_layernorm_affine_grad(cuda_device, rows, c):
    ops = _ops()
    x = (rnd(rows, c, scale=2.0, seed=5) - 0.3).to(BF16)
    dy = rnd(rows, c, seed=6).to(BF16)
    dg, db = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
    ops.layernorm_affine_grad(x, dy, dg, db, 1e-5)
    rg, rb = torch.ones(c), torch.zeros(c)
    _mock().layernorm_affine_grad(x.float().cpu(), dy.float().cpu(), rg, rb, 1e-5)
    assert_close(dg.cpu(), rg, rtol=2e-3, atol_scale=1e-3, what=f"ln dgamma rows={rows} c={c}")
    assert_close(db.cpu(), rb, rtol=2e-3, atol_scale=1e-3, what=f"ln dbeta rows={rows} c={c}")


@pytest.mark.parametrize("n,offset", [(1 << 20, 0), (1000003, 0), (4099, 1), (3, 0)])
def test_ema_update(cuda_device, n, offset):
    ops = _ops()
    tgt = rnd(n + offset, seed=7)[offset:]           # offset 1: a 4-byte-aligned slice takes the scalar path
    src = rnd(n + offset, seed=8)[offset:]
    ref = tgt.clone().mul_(0.95).add_(src, alpha=0.05)
    ops.ema_update(tgt, src, 0.95)
    assert torch.allclose(tgt, ref, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("kind,shape,cout", [("linear", (4096, 320), 320), ("linear", (616, 1024), 640), ("linear", (2, 256), 320),
                                             ("conv2d", (4, 16, 24, 128), 192), ("conv3d", (2, 8, 160, 64), 128)])
def test_wgrad_wide_base_weight_gradient(cuda_device, kind, shape, cout):
    """dW of a whole layer through 64-column slices of dy read in place == the contract (and == wgrad on copied slices)."""
    ops = _ops()
    x = rnd(*shape, seed=9).to(BF16)
    dy = rnd(*shape[:-1], cout, seed=10).to(BF16)
    cin = shape[-1]
    taps = None if kind == "linear" else (ops._TAPS_3X3 if kind == "conv2d" else ops._TAPS_T3)
    nt = 1 if taps is None else len(taps)
    out = torch.zeros(cout, cin, nt, device="cuda")
    ops.wgrad_wide(x, dy, out, taps=taps, out_strides=(cin * nt, nt, 1))
    ref = torch.zeros(cout, cin, nt)
    _mock().wgrad_wide(x.float().cpu(), dy.float().cpu(), ref, taps=taps, out_strides=(cin * nt, nt, 1))
    assert_close(out.cpu(), ref, rtol=4e-3, atol_scale=2e-3, what=f"wgrad_wide {kind} {shape} -> {cout}")
    if shape[0] >= 16:
        out2 = torch.zeros_like(out)
        for j0 in range(0, cout, 64):
            ops.wgrad(x, dy[..., j0:j0 + 64].contiguous(), out2[j0:j0 + 64], taps=taps, out_strides=(cin * nt, nt, 1))
        assert_close(out, out2, rtol=1e-3, atol_scale=1e-4, what="wgrad_wide vs wgrad on copied slices")   # fp32 atomics: order only


# ----------------------------------------------------------------------------- the whole step
def _setup(with_ema=True):
    from oracle.configs import UNET_CONFIGS
    from oracle.weights import seeded_state_dict
    from t2v_turbo_b200.distill_v2 import V2Step
    from t2v_turbo_b200.full_train import FullUNet
    from t2v_turbo_b200.scheduler import T2VTurboScheduler
    from t2v_turbo_b200.unet import UNetModel
    g = torch.load(os.path.join(GOLD, "v2_step_small_motion.pt"))
    spec = UNET_CONFIGS["small_motion"]
    m = UNetModel(**spec["cfg"])
    sd = seeded_state_dict(m.state_dict(), spec["weight_seed"])
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    s = FullUNet(m, with_target=with_ema).eval()
    s.pack()
    target = None
    if with_ema:
        gq = torch.Generator().manual_seed(g["target_perturb_seed"])
        for n, p in m.named_parameters():
            s.arena.view(s.arena.target, s.arena.index[n]).copy_(sd[n] * (1.0 + 0.02 * torch.randn(p.shape, generator=gq)))
        target = UNetModel(**spec["cfg"])
        target.load_state_dict(sd, strict=True)
        target = target.cuda().eval()
        s.arena.bind(target, s.arena.target)           # the EMA network's parameters ARE the target arena
        target.invalidate_packed()
    h = g["hyper"]
    step = V2Step(s, T2VTurboScheduler(linear_start=0.00085, linear_end=0.012), target_unet=target, num_ddim_timesteps=h["n_ddim"],
                  topk=h["topk"], motion_gs=h["motion_gs"], percentage=h["percentage"], use_motion_cond=True,
                  timestep_scaling_factor=h["ts_scale"])
    return g, s, step, sd


def test_v2_step_vs_reference_composition(cuda_device):
    from t2v_turbo_b200.distill_v2 import train_step_v2
    g, s, step, sd = _setup(with_ema=True)
    inp, h = g["inputs"], g["hyper"]
    batch = {k: inp[k].cuda() for k in ("index", "z_t", "cond_teacher_out", "uncond_teacher_out", "score", "use_motion_guide", "prompt_emb")}
    p0 = s.arena.params.clone()
    out = train_step_v2(step, batch, lr=h["lr"], temporal_lr_scale=h["temporal_lr_scale"], ema_decay=h["ema_decay"],
                        max_grad_norm=h["max_grad_norm"], weight_decay=h["weight_decay"], fixed=dict(w=inp["w"]))
    torch.cuda.synchronize()
    e = {k: _rel(out[k], g[k]) for k in ("model_pred", "x_prev", "target")}
    loss, loss_ref = float(out["loss"]), float(g["loss"])
    print(f"\n[v2 small] loss {loss:.6f} vs reference {loss_ref:.6f}; rel-L2 {e}")
    assert e["x_prev"] < 1e-5 and e["model_pred"] < 3e-2 and e["target"] < 3e-2, e          # x_prev involves no network
    assert abs(loss - loss_ref) < 3e-2 * loss_ref, (loss, loss_ref)
    names = s.arena.names
    assert torch.isfinite(s.arena.grads).all()
    ratio = torch.tensor([s.arena.grad(n).double().norm().item() / max(g["grad_norms"][n], 1e-30) for n in names])
    print(f"[v2 small] grad-norm ratio ours/reference: min {ratio.min():.4f} max {ratio.max():.4f} over {len(names)} tensors")
    rels = {n: _rel(s.arena.grad(n), sc * t.float()) for n, (sc, t) in g["grads_full"].items()}
    wn = max(rels, key=rels.get)
    print(f"[v2 small] full-tensor rel-L2 over {len(rels)} tensors: median {sorted(rels.values())[len(rels) // 2]:.3e}, worst {rels[wn]:.3e} ({wn})")
    assert (ratio - 1).abs().max() < 8e-2 and rels[wn] < 1.2e-1, (ratio.min(), ratio.max(), wn, rels[wn])
    assert abs(float(s.arena.grad_norm()) - g["total_norm"]) < 3e-2 * g["total_norm"]
    # optimizer + EMA: the first AdamW step moves each weight by ~lr * sign(grad): compare where the reference gradient is not noise-level
    moved = (s.arena.params - p0).abs()
    tmask = torch.zeros_like(moved, dtype=torch.bool)
    for lo, hi, temporal in s.arena.runs:
        tmask[lo:hi] = temporal
    bound = h["lr"] * (1 + h["weight_decay"] * float(p0.abs().max()))
    assert float(moved[~tmask].max()) <= 1.05 * bound + 1e-9 and float(moved[tmask].max()) <= 1.05 * bound * h["temporal_lr_scale"] + 1e-9
    assert float(moved[tmask].mean()) > 2.0 * float(moved[~tmask].mean()), "temporal group must step with lr * temporal_lr_scale"
    for n, t in g["ema_after"].items():
        assert _rel(s.arena.view(s.arena.target, s.arena.index[n]), t) < 1e-4, n


def test_v2_step_self_target_and_second_step(cuda_device):
    from t2v_turbo_b200.distill_v2 import train_step_v2
    g, s, step, _ = _setup(with_ema=False)
    inp = g["inputs"]
    batch = {k: inp[k].cuda() for k in ("index", "z_t", "cond_teacher_out", "uncond_teacher_out", "score", "prompt_emb")}
    s.arena.zero_grad()
    out = step(batch, fixed=dict(w=inp["w"]))
    assert _rel(out["target"], g["target_self"]) < 3e-2
    assert abs(float(out["loss"]) - float(g["loss_self_target"])) < 3e-2 * float(g["loss_self_target"])
    s.train()                                        # dropouts on: finite gradients, and a step changes the prediction
    out1 = train_step_v2(step, batch, lr=1e-4, fixed=dict(w=inp["w"]))
    assert torch.isfinite(s.arena.grads).all() and float(s.arena.grad_norm()) > 0
    s.eval()
    s.arena.zero_grad()
    out2 = step(batch, fixed=dict(w=inp["w"]))
    assert _rel(out2["model_pred"], out["model_pred"]) > 1e-5, "the optimizer step did not change the student's prediction"
    assert torch.isfinite(out1["loss"]).all()
