"""GPU parity of the student UNet's training forward AND backward (t2v_turbo_b200/train_unet.py) against the UNMODIFIED reference:
tests/golden/student_grads_small.pt holds the output and the LoRA gradients of the reference UNet with LoRA injected by the
reference's own `inject_trainable_lora_extended` (fp32 autograd on CPU; oracle/make_goldens.py::gen_student_grads).

Tolerances: bf16 activations AND bf16 gradients through ~50 layers against fp32 autograd.  Bounds are <= 2x the error observed
on B200 (printed by the test)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _rel(got, ref):
    got, ref = got.double().cpu(), ref.double().cpu()
    return ((got - ref).norm() / (ref.norm() + 1e-30)).item()


def _student(name="small"):
    from oracle.configs import UNET_CONFIGS, student_loras, unet_inputs
    from oracle.weights import seeded_state_dict
    from t2v_turbo_b200.train_unet import StudentUNet
    from t2v_turbo_b200.unet import UNetModel
    spec = UNET_CONFIGS[name]
    m = UNetModel(**spec["cfg"])
    m.load_state_dict(seeded_state_dict(m.state_dict(), spec["weight_seed"]), strict=True)
    m = m.cuda().eval()
    s = StudentUNet(m, r=64, dropout_p=0.1, scale=1.0).eval()      # eval: every dropout off, like the fixture
    return spec, m, s, student_loras, unet_inputs


def test_student_forward_backward_vs_reference_autograd(cuda_device):
    g = torch.load(os.path.join(GOLD, "student_grads_small.pt"))
    spec, m, s, student_loras, unet_inputs = _student()
    assert [tuple(x) for x in g["shapes"]] == s.arena.shapes, "LoRA list layout differs from the reference's flat list"
    s.arena.load_list(student_loras(g["shapes"]))
    s.pack()
    inp = unet_inputs(spec, g["timestep"])
    y = s(inp["x"].cuda(), inp["timesteps"].cuda(), context=inp["context"].cuda(), fps=16, timestep_cond=inp["timestep_cond"].cuda())
    e_y = _rel(y, g["output"])
    print(f"\n[student small] forward rel-L2 vs reference {e_y:.3e}")
    assert e_y < 3e-2, e_y
    s.arena.zero_grad()
    s.backward(g["d_out"].cuda())
    torch.cuda.synchronize()
    norms = g["grad_norms"]
    n = len(s.arena.shapes)
    ours = [s.arena.grad(i) for i in range(n)]
    assert all(torch.isfinite(t).all() for t in ours)
    # (1) every gradient's norm
    ratio = torch.tensor([ours[i].double().norm().item() / max(norms[i].item(), 1e-30) for i in range(n)])
    worst = (ratio - 1).abs().max().item()
    print(f"[student small] grad-norm ratio ours/reference: min {ratio.min():.4f} max {ratio.max():.4f} over {n} tensors")
    # (2) the stored layers in full
    rels = {j: _rel(ours[j], sc * t.float()) for j, (sc, t) in g["grads_full"].items()}
    wj = max(rels, key=rels.get)
    print(f"[student small] full-tensor rel-L2 over {len(rels)} tensors: median {sorted(rels.values())[len(rels) // 2]:.3e}, "
          f"worst {rels[wj]:.3e} (tensor {wj}, shape {s.arena.shapes[wj]}, layer {s.layer_list[wj // 2].name})")
    total = _rel(torch.cat([ours[j].flatten() for j in rels]), torch.cat([(sc * t.float()).flatten() for sc, t in g["grads_full"].values()]))
    print(f"[student small] stored tensors concatenated rel-L2 {total:.3e}")
    # Yardstick: the unmodified reference run in bf16 (weights, activations, autograd) against its own fp32 gradients, stored in
    # the fixture: output 2.2e-2, gradients median 4.0e-2 / worst 6.0e-2 / concatenated 3.8e-2, norm ratio 0.975 .. 1.021.
    # Observed on B200: output 1.9e-2, median 3.6e-2, worst 5.2e-2, concatenated 3.2e-2, norm ratio 0.985 .. 1.011.
    rb = g["ref_bf16"]
    print(f"[student small] reference bf16 vs its own fp32: {rb}")
    assert e_y <= 1.15 * rb["output_rel"], (e_y, rb["output_rel"])
    assert worst < 3e-2, f"gradient norm off by {worst:.3f}"
    assert rels[wj] <= 1.3 * rb["grad_rel_worst"], (wj, rels[wj], rb["grad_rel_worst"])
    assert total <= 1.15 * rb["grad_rel_concat"], (total, rb["grad_rel_concat"])


def test_student_training_mode_and_optimizer_step(cuda_device):
    """Training mode (LoRA + temporal-conv dropouts on): finite gradients, a fused AdamW step changes the output, and the
    reference's initial state (lora_up = 0) reproduces the frozen UNet's own forward."""
    spec, m, s, student_loras, unet_inputs = _student()
    s.pack()
    inp = unet_inputs(spec, 519)
    args = (inp["x"].cuda(), inp["timesteps"].cuda())
    kw = dict(context=inp["context"].cuda(), fps=16, timestep_cond=inp["timestep_cond"].cuda())
    y0 = s(*args, **kw)
    base = m(*args, **kw)
    assert _rel(y0, base) < 2.5e-2, _rel(y0, base)          # up = 0: the LoRA branch is exactly zero; two bf16 paths of one model
    s.train()
    torch.manual_seed(0)
    y = s(*args, **kw)
    s.arena.zero_grad()
    s.backward(torch.randn_like(y))
    assert torch.isfinite(s.arena.grads).all()
    gn = float(s.arena.grad_norm())
    assert gn > 0
    # up = 0 => d(lora_down) = 0 and d(lora_up) != 0 (utils/lora.py:40-43 initialisation)
    assert float(s.arena.grad(1).abs().max()) == 0.0 and float(s.arena.grad(0).abs().max()) > 0.0
    s.arena.adamw_step(lr=1e-3, max_grad_norm=1.0)
    s.refresh()
    s.eval()
    y1 = s(*args, **kw)
    assert _rel(y1, y0) > 1e-4, "the optimizer step did not change the student's output"


def test_distill_step_vs_reference_composition(cuda_device):
    """One consistency-distillation step (DistillStep: add_noise, student, teacher CFG + DDIM, target, pseudo-Huber loss, student
    backward) against the same step composed from the UNMODIFIED reference's pieces in fp32 (gen_distill_step), same draws."""
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
    teacher = UNetModel(**tcfg)
    teacher.load_state_dict({k: v for k, v in sd.items() if not k.startswith("time_cond_proj")}, strict=True)
    base, teacher = base.cuda().eval(), teacher.cuda().eval()
    s = StudentUNet(base, r=64).eval()
    assert [tuple(x) for x in g["shapes"]] == s.arena.shapes
    s.arena.load_list(student_loras(g["shapes"]))
    s.pack()
    step = DistillStep(s, teacher, T2VTurboScheduler(linear_start=0.00085, linear_end=0.012), num_ddim_timesteps=50, topk=20,
                       loss_type="huber", huber_c=0.001, timestep_scaling_factor=10.0)
    inp = g["inputs"]
    s.arena.zero_grad()
    out = step(inp["latents"].cuda(), inp["prompt"].cuda(), inp["uncond"].cuda(),
               fixed=dict(index=inp["index"], noise=inp["noise"].cuda(), w=inp["w"]))
    torch.cuda.synchronize()
    assert out["start_timesteps"].tolist() == g["start_timesteps"].tolist() and out["timesteps"].tolist() == g["timesteps"].tolist()
    e = {k: _rel(out[k], g[k]) for k in ("model_pred", "x_prev", "target")}
    loss, loss_ref = float(out["loss"]), float(g["loss"])
    print(f"\n[distill small] loss {loss:.6f} vs reference {loss_ref:.6f}; rel-L2 {e}")
    assert e["x_prev"] < 1.5e-2 and e["model_pred"] < 2.5e-2 and e["target"] < 2.5e-2, e
    assert abs(loss - loss_ref) < 3e-2 * loss_ref, (loss, loss_ref)
    n = len(s.arena.shapes)
    ours = [s.arena.grad(i) for i in range(n)]
    ratio = torch.tensor([ours[i].double().norm().item() / max(g["grad_norms"][i].item(), 1e-30) for i in range(n)])
    rels = {j: _rel(ours[j], sc * t.float()) for j, (sc, t) in g["grads_full"].items()}
    total = _rel(torch.cat([ours[j].flatten() for j in rels]), torch.cat([(sc * t.float()).flatten() for sc, t in g["grads_full"].values()]))
    print(f"[distill small] grad-norm ratio min {ratio.min():.4f} max {ratio.max():.4f}; stored tensors rel-L2 median "
          f"{sorted(rels.values())[len(rels) // 2]:.3e} worst {max(rels.values()):.3e} concatenated {total:.3e}")
    # The loss gradient is sign-like (d / sqrt(d^2 + c^2) with c = 1e-3 and |d| ~ 0.16) in d = model_pred - target, a difference
    # of two predictions that each carry ~2e-2 of bf16 error: roughly one element in ten has |d| inside that noise and may flip,
    # which bounds the agreement of ANY bf16 run with the fp32 fixture at ~0.3 rel-L2.  Observed on B200 with two equally valid
    # roundings of the same backward: 0.287 / norm ratios 0.92 .. 1.10 (dx = dy W, then += dt D) and 0.333 / 0.90 .. 1.17 (one
    # GEMM over the (dy | dt) pair) — the measure is noise-dominated, so its bound is a sanity bound, not 2x an observation.
    # The backward itself is pinned by the linear-loss fixture above (3.2e-2); this test pins the step's glue: timesteps,
    # add_noise, the CFG / DDIM algebra, the boundary scalings, the loss value, and the gradient's direction.
    assert (ratio - 1).abs().max().item() < 0.25 and total < 0.45, (ratio.min().item(), ratio.max().item(), total)


def test_graphed_distill_step_matches_eager(cuda_device):
    """GraphedDistillStep (the device side of the step as a chain of CUDA graphs cut at the gradient-arena hooks) reproduces the
    eager step: same loss and gradients for the same draws (eval mode: no dropout randomness), replayed twice with different
    draws in between (the static buffers really are re-read), and the cuts report ascending-completion offsets."""
    from oracle.configs import UNET_CONFIGS, student_loras
    from oracle.weights import seeded_state_dict
    from t2v_turbo_b200.distill import DistillStep, GraphedDistillStep
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
    teacher = UNetModel(**tcfg)
    teacher.load_state_dict({k: v for k, v in sd.items() if not k.startswith("time_cond_proj")}, strict=True)
    base, teacher = base.cuda().eval(), teacher.cuda().eval()
    s = StudentUNet(base, r=64).eval()
    s.arena.load_list(student_loras(g["shapes"]))
    s.pack()
    step = DistillStep(s, teacher, T2VTurboScheduler(linear_start=0.00085, linear_end=0.012))
    inp = g["inputs"]
    lat, pr, un = inp["latents"].cuda(), inp["prompt"].cuda(), inp["uncond"].cuda()
    fixed = dict(index=inp["index"], noise=inp["noise"].cuda(), w=inp["w"])
    s.arena.zero_grad()
    out_e = step(lat, pr, un, fixed=fixed)
    grads_e, loss_e = s.arena.grads.clone(), float(out_e["loss"])
    pred_e, tgt_e = out_e["model_pred"].clone(), out_e["target"].clone()
    seen = []

    class Rec:
        def ready(self, off):
            seen.append(off)
    gs = GraphedDistillStep(step, lat, pr, un, reducer=Rec())
    assert len(gs.segments) >= 4
    for rep in range(2):
        s.arena.zero_grad()
        gs(lat, pr, un, fixed=dict(index=torch.tensor([3, 44]), noise=torch.randn_like(lat), w=torch.tensor([9.0, 5.5])))   # other draws
        s.arena.zero_grad()
        seen.clear()
        out_g = gs(lat, pr, un, fixed=fixed)
        torch.cuda.synchronize()
        # Not bit-equal by construction: GroupNorm statistics and split-K partial sums are fp32 atomics (order varies run to
        # run), a last-bit difference that bf16 re-rounding amplifies to ~1e-3 in the predictions (two EAGER runs differ as
        # much: observed losses 0.1599 / 0.1616 for this fixture); the loss gradient is sign-like, see the test above.
        e_p, e_t = _rel(out_g["model_pred"], pred_e), _rel(out_g["target"], tgt_e)
        e_g = _rel(s.arena.grads, grads_e)
        print(f"\n[graphed distill] rep {rep}: loss {float(out_g['loss']):.6f} vs eager {loss_e:.6f}; model_pred {e_p:.2e}, target {e_t:.2e}, grads {e_g:.3f}")
        # observed: model_pred 5.7e-3, target 1.1e-2 between the graphed and the eager run; against the REFERENCE fixture both
        # sit at the same distance (7.6e-3 / 8.1e-3 eager), which is the meaningful check
        assert e_p < 2e-2 and e_t < 2.5e-2, (e_p, e_t)
        assert _rel(out_g["model_pred"], g["model_pred"]) < 2.5e-2 and _rel(out_g["target"], g["target"]) < 2.5e-2
        assert abs(float(out_g["loss"]) - loss_e) < 3e-2 * abs(loss_e), (float(out_g["loss"]), loss_e)
        assert e_g < 0.4, e_g
        assert seen == sorted(seen, reverse=True) and seen[-1] == 0, seen
    # the in-place operand refresh under a captured graph: an optimizer step changes what the SAME graphs compute
    s.graph_refresh()
    s.arena.adamw_step(lr=3e-2, max_grad_norm=1.0)
    s.refresh()
    s.arena.zero_grad()
    out2 = gs(lat, pr, un, fixed=fixed)
    assert _rel(out2["model_pred"], pred_e) > 2e-2, "the captured graphs did not see the refreshed LoRA operands"
