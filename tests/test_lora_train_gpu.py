"""GPU parity tests of the LoRA training ops (SURVEY §8 a23 / e): the LoRA-injected layers' forward and backward, the
weight-gradient kernel, the arena, fused AdamW, gradient norm and the MSE loss — against goldens produced by the
UNMODIFIED reference modules (utils/lora.py:19-230, autograd) and against torch's own AdamW / mse_loss."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
BF16 = torch.bfloat16


def _rel(got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    return ((got - ref).norm() / ref.norm()).item()


def _make(kind, g):
    from t2v_turbo_b200 import lora_train as lt
    w = g["w"]
    if kind == "linear":
        m = lt.LoraInjectedLinear(w.shape[1], w.shape[0], bias=g["b"] is not None, r=64, dropout_p=0.0, scale=g["scale"])
        base = m.linear
    elif kind == "conv2d":
        m = lt.LoraInjectedConv2d(w.shape[1], w.shape[0], 3, padding=1, r=64, dropout_p=0.0, scale=g["scale"])
        base = m.conv
    else:
        m = lt.LoraInjectedConv3d(w.shape[1], w.shape[0], (3, 1, 1), padding=(1, 0, 0), bias=g["b"] is not None, r=64, dropout_p=0.0,
                                  scale=g["scale"])
        base = m.conv
    with torch.no_grad():
        base.weight.copy_(w)
        if g["b"] is not None:
            base.bias.copy_(g["b"])
        m.lora_up.weight.copy_(g["up"])
        m.lora_down.weight.copy_(g["down"])
    return m.cuda()


@pytest.mark.parametrize("kind", ["linear", "conv2d", "conv3d"])
def test_lora_layer_forward_backward_vs_reference(cuda_device, kind):
    """y, dx, d lora_up, d lora_down of LoraInjectedLinear / Conv2d / Conv3d against the reference module's autograd
    (fp32 on CPU); ours computes in bf16 with fp32 accumulation.  Observed rel-L2 is printed; bounds are ~2x observed."""
    g = torch.load(os.path.join(GOLD, "lora_layers.pt"))[kind]
    m = _make(kind, g)
    x = g["x"].cuda().requires_grad_(True)
    y = m(x)
    y.backward(g["dy"].cuda())
    errs = dict(y=_rel(y, g["y"]), dx=_rel(x.grad, g["dx"]), d_up=_rel(m.lora_up.weight.grad, g["d_up"]),
                d_down=_rel(m.lora_down.weight.grad, g["d_down"]))
    print(f"\n[lora {kind}] rel-L2 vs reference autograd: {errs}")
    assert m.lora_up.weight.grad.dtype == torch.float32
    for k, v in errs.items():
        assert v <= 1.2e-2, (k, errs)
    # gradients ACCUMULATE (the arena is zeroed once per optimizer step): a second backward doubles them
    y2 = m(x.detach())
    y2.backward(g["dy"].cuda())
    assert _rel(m.lora_up.weight.grad, 2 * g["d_up"]) <= 1.2e-2 and _rel(m.lora_down.weight.grad, 2 * g["d_down"]) <= 1.2e-2


def test_lora_dropout_mask_is_shared_by_forward_and_backward(cuda_device):
    """Training mode: the keep-mask drawn in the forward scales the LoRA branch by 1 / (1 - p) and is reused in the backward
    (checked against the same computation done with the mask in fp32 torch)."""
    from t2v_turbo_b200 import lora_train as lt
    torch.manual_seed(5)
    m = lt.LoraInjectedLinear(128, 192, bias=False, r=64, dropout_p=0.25, scale=1.0).cuda().train()
    with torch.no_grad():
        m.lora_up.weight.normal_(0, 0.05)
    x = torch.randn(256, 128, device="cuda")
    pk = m._packed()
    mask, ms = lt._keep_mask((256, 192), 0.25, x.device)
    assert 0.70 <= mask.float().mean().item() <= 0.80 and abs(ms - 1 / 0.75) < 1e-6
    y, t, _, _ = lt.lora_forward(pk, x.to(BF16), mask, ms)
    xb = x.to(BF16).float()
    ref = xb @ m.linear.weight.float().t() + (xb @ m.lora_down.weight.float().t() @ m.lora_up.weight.float().t()) * mask.float() * ms
    assert _rel(y, ref) <= 1e-2
    gu, gd = torch.zeros_like(m.lora_up.weight), torch.zeros_like(m.lora_down.weight)
    dy = torch.randn(256, 192, device="cuda").to(BF16)
    dx = lt.lora_backward(pk, x.to(BF16), t, mask, ms, dy, gu, gd)
    du = dy.float() * mask.float() * ms
    assert _rel(gu, du.t() @ (xb @ m.lora_down.weight.float().t())) <= 1.5e-2
    assert _rel(gd, (du @ m.lora_up.weight.float()).t() @ xb) <= 1.5e-2
    assert _rel(dx, dy.float() @ m.linear.weight.float() + (du @ m.lora_up.weight.float()) @ m.lora_down.weight.float()) <= 1.5e-2


def test_dropout_scale_kernel(cuda_device):
    """t2v_dropout_scale: out = x * scale / (1 - p) * keep with keep ~ Bernoulli(1 - p) drawn in the kernel (Philox4x32-10) and
    written for the adjoint.  Checks: out is exactly x * scale' * keep (bf16 rounding of one product), the keep rate and its
    independence across positions / calls, reproducibility for a fixed (seed, call id), fresh masks after dropout_advance."""
    from t2v_turbo_b200 import ops
    p, scale = 0.1, 0.5
    x = torch.randn(4096, 640, device="cuda").to(BF16)
    ops.dropout_seed(x.device, seed=1234)
    y, keep = ops.dropout_scale(x, p, scale)
    assert keep.dtype == torch.uint8 and keep.shape == x.shape and set(keep.unique().tolist()) <= {0, 1}
    ref = (x.float() * (scale / (1 - p)) * keep.float()).to(BF16)
    assert torch.equal(y, ref)
    n = keep.numel()
    rate = keep.float().mean().item()
    sigma = (p * (1 - p) / n) ** 0.5
    assert abs(rate - (1 - p)) < 5 * sigma, (rate, sigma)                     # keep rate exact to sampling noise
    k = keep.float() - (1 - p)
    for shift in (1, 7, 8, 640):                                              # no correlation between neighbouring draws
        c = (k.flatten()[:-shift] * k.flatten()[shift:]).mean().item() / (p * (1 - p))
        assert abs(c) < 5 / n ** 0.5, (shift, c)
    col = keep.float().mean(0)
    assert (col - (1 - p)).abs().max().item() < 6 * (p * (1 - p) / keep.shape[0]) ** 0.5   # every column sees the same rate
    # a second call (next call id) draws an independent mask; the same (seed, call id) reproduces the first one
    y2, keep2 = ops.dropout_scale(x, p, scale)
    agree = (keep2 == keep).float().mean().item()
    assert abs(agree - ((1 - p) ** 2 + p ** 2)) < 5e-3, agree
    ops._DROPOUT_CALLS[0] -= 2
    y3, keep3 = ops.dropout_scale(x, p, scale)
    assert torch.equal(keep3, keep) and torch.equal(y3, y)
    ops._DROPOUT_CALLS[0] -= 1
    ops.dropout_advance(x.device)                                             # next step: same call id, new seed
    _, keep4 = ops.dropout_scale(x, p, scale)
    assert abs((keep4 == keep).float().mean().item() - ((1 - p) ** 2 + p ** 2)) < 5e-3
    # the residual add that follows the layer folded into the same pass (fp32 add before the one bf16 rounding)
    ops._DROPOUT_CALLS[0] -= 1
    a = torch.randn_like(x)
    y5, keep5 = ops.dropout_scale(x, p, scale, addend=a)
    assert torch.equal(keep5, keep4)
    ref5 = x.float() * (scale / (1 - p)) * keep5.float() + a.float()
    assert (y5.float() - ref5).abs().max().item() <= 2.0 ** -8 * ref5.abs().max().item()
    assert ((y5.float() - ref5).norm() / ref5.norm()).item() < 3e-3
    # the adjoint reuses the stored mask
    dy = torch.randn_like(x)
    dx = ops.scale_mask(dy, scale / (1 - p), keep)
    assert torch.equal(dx, (dy.float() * (scale / (1 - p)) * keep.float()).to(BF16))


@pytest.mark.parametrize("shape,p,seed", [((4096, 320), 0.1, 1234), ((24, 64), 0.5, 2 ** 40 + 17), ((1000, 1280), 0.25, 99)])
def test_dropout_mask_bit_exact_vs_philox_oracle(cuda_device, shape, p, seed):
    """The keep-mask drawn inside t2v_dropout_scale equals, bit for bit, the mask the numpy Philox4x32-10 oracle draws for the same
    (seed, call id) — the oracle is pinned on Random123's known-answer vectors in the CPU suite (tests/test_oracle.py)."""
    import numpy as np
    from oracle.philox_oracle import keep_mask
    from t2v_turbo_b200 import ops
    x = torch.randn(*shape, device="cuda").to(BF16)
    ops.dropout_seed(x.device, seed=seed)
    _, keep = ops.dropout_scale(x, p)
    want = keep_mask(x.numel(), 1.0 - p, seed, ops._DROPOUT_CALLS[0]).reshape(shape)
    got = keep.cpu().numpy()
    assert np.array_equal(got, want), f"{int((got != want).sum())} of {got.size} mask bytes differ"


def test_dropout_scale_fresh_masks_under_cuda_graph_replay(cuda_device):
    """The seed lives on the device and is advanced by a captured add: every replay of a captured step draws a new mask."""
    from t2v_turbo_b200 import ops
    x = torch.randn(512, 256, device="cuda").to(BF16)
    ops.dropout_seed(x.device, seed=7)
    ops.dropout_scale(x, 0.5)                                                 # warm-up (lazy state) before capture
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        ops.dropout_advance(x.device)
        _, keep = ops.dropout_scale(x, 0.5)
    masks = []
    for _ in range(3):
        g.replay()
        torch.cuda.synchronize()
        masks.append(keep.clone())
    for a in range(3):
        assert abs(masks[a].float().mean().item() - 0.5) < 0.01
        for b in range(a + 1, 3):
            assert abs((masks[a] == masks[b]).float().mean().item() - 0.5) < 0.01


@pytest.mark.parametrize("pts,c,r,taps", [((1000,), 320, 64, None), ((2, 20, 32), 128, 64, "3x3"), ((1, 16, 160), 64, 64, "t3"),
                                           ((5,), 1280, 64, None), ((3, 8, 8), 192, 32, "3x3")])
def test_wgrad_kernel(cuda_device, pts, c, r, taps):
    """t2v_wgrad (MN-major operands in place, taps as TMA coordinate offsets, zero fill = conv padding) vs fp32 torch."""
    from t2v_turbo_b200 import ops
    g = torch.Generator().manual_seed(9)
    a = torch.randn(*pts, c, generator=g).to(BF16).cuda()
    b = torch.randn(*pts, r, generator=g).to(BF16).cuda()
    tl = None if taps is None else (ops._TAPS_3X3 if taps == "3x3" else ops._TAPS_T3)
    nt = 1 if tl is None else len(tl)
    out = torch.full((r, c, nt), 0.5, device="cuda")
    ops.wgrad(a, b, out, taps=tl, out_strides=(c * nt, nt, 1), alpha=0.5)
    af, bf = a.float(), b.float()
    if taps is None:
        ref = torch.einsum("pc,pj->jc", af.view(-1, c), bf.view(-1, r))[:, :, None]
    elif taps == "3x3":   # a: [n, h, w, c]
        ap = torch.nn.functional.pad(af, (0, 0, 1, 1, 1, 1))
        n, h, w = pts
        ref = torch.stack([torch.einsum("nhwc,nhwj->jc", ap[:, ky:ky + h, kx:kx + w], bf) for ky in range(3) for kx in range(3)], -1)
    else:                 # a: [b, t, hw, c]
        ap = torch.nn.functional.pad(af, (0, 0, 0, 0, 1, 1))
        t = pts[1]
        ref = torch.stack([torch.einsum("bthc,bthj->jc", ap[:, kt:kt + t], bf) for kt in range(3)], -1)
    ref = 0.5 + 0.5 * ref
    e = _rel(out, ref)
    print(f"\n[wgrad pts={pts} c={c} r={r} taps={taps}] rel-L2 {e:.2e}")
    assert e <= 2e-3


def test_arena_adamw_grad_norm_and_mse(cuda_device):
    from t2v_turbo_b200 import ops
    from t2v_turbo_b200.lora_train import LoraArena
    shapes = [(320, 64), (64, 320), (640, 64, 1, 1), (64, 640, 3, 3), (7, 3)]
    arena = LoraArena(shapes, "cuda")
    assert arena.numel == sum(torch.Size(s).numel() for s in shapes) and arena.padded % 4 == 0
    g = torch.Generator().manual_seed(3)
    ps = [torch.randn(s, generator=g) for s in shapes]
    arena.load_list(ps)
    ref_params = [torch.nn.Parameter(p.clone().cuda()) for p in ps]
    opt = torch.optim.AdamW(ref_params, lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    for it in range(3):
        grads = [torch.randn(s, generator=g) * 3 for s in shapes]
        arena.zero_grad()
        for i, gr in enumerate(grads):
            arena.grad(i).copy_(gr)
            ref_params[i].grad = gr.clone().cuda() / 2.0            # DDP mean over 2 ranks
        ref_norm = torch.nn.utils.clip_grad_norm_(ref_params, 1.0)
        got_norm = arena.grad_norm(grad_scale=0.5)
        torch.testing.assert_close(got_norm.item(), ref_norm.item(), rtol=1e-5, atol=1e-6)
        opt.step()
        arena.adamw_step(lr=1e-2, grad_scale=0.5, max_grad_norm=1.0)
    for i, p in enumerate(ref_params):
        torch.testing.assert_close(arena.param(i), p.data, rtol=2e-5, atol=2e-6)
    back = arena.to_list()
    assert [tuple(t.shape) for t in back] == shapes
    a = torch.randn(2, 4, 16, 40, 64, device="cuda").to(BF16)
    b = torch.randn(2, 4, 16, 40, 64, device="cuda").to(BF16)
    loss, grad = ops.mse_loss_grad(a, b)
    ref = torch.nn.functional.mse_loss(a.float(), b.float())
    torch.testing.assert_close(loss.item(), ref.item(), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(grad.float(), (2 * (a.float() - b.float()) / a.numel()).to(BF16).float(), rtol=1e-2, atol=1e-9)
