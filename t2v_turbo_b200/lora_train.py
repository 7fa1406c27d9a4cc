"""LoRA-injected layers for the consistency-distillation step on B200: forward AND backward through the C ABI.

Mirrors `utils/lora.py:19-230` (`LoraInjectedLinear / Conv2d / Conv3d`):

    y = base(x) + dropout(lora_up(lora_down(x))) * scale          r = min(r, in, out); lora_down has the base kernel,
                                                                  lora_up is 1x1; down ~ N(0, 1/r), up = 0

with the base layer frozen and only `lora_up / lora_down` trained (`train_t2v_turbo_v1_lora.py:862-906`).  Everything
arithmetic is a libt2v_b200.so kernel:

  forward   t = down(x) -> u = up(t) -> branch = u * mask * scale / (1 - p) -> y = base(x) + branch     3 tcgen05 GEMMs
            (the base GEMM takes `branch` as its residual operand: one pass over y)
  backward  du = dy * mask * scale / (1 - p);  dt = du U;  dx = (dy | dt) [W^T | D^T]                     2 tcgen05 GEMMs
            dU += du^T t,  dD += dt^T x (per tap)        t2v_wgrad: MN-major operands read in place, fp32 red.add
            straight into the GRADIENT ARENA — one contiguous fp32 buffer for all layers (117 142 176 values for
            the VC2 UNet at r = 64), which `dist.allreduce_arena` reduces with ONE NCCL all-reduce and
            `LoraArena.adamw_step` updates with ONE fused AdamW launch.

dgrad needs no kernel of its own: dx = dy W is the forward implicit GEMM on transposed (and, for convolutions,
tap-reversed) weights packed once at load.  Activations are channels-last bf16 like the inference path; the modules
accept the reference's layouts ([..., K] for Linear, NCHW / NCDHW for the convolutions) and convert at the boundary.
The dropout keep-mask is drawn INSIDE the scale kernel (`t2v_dropout_scale`: Philox4x32-10 keyed by a device-resident seed
that `ops.dropout_advance` moves between steps — the same distribution as nn.Dropout, not torch's random stream) and kept as
one byte per element for the adjoint.

The layers between the LoRA layers (GroupNorm, LayerNorm, attention, GEGLU, SiLU, the strided / upsampling convolutions)
and the whole-UNet traversal live in `train_unet.StudentUNet`; this module is the LoRA layer pair itself, the gradient
arena, and the reference-named `LoraInjected*` modules (autograd glue for use outside StudentUNet).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import ops

BF16 = torch.bfloat16
_TAPS_3X3 = ops._TAPS_3X3
_TAPS_T3 = ops._TAPS_T3


# =============================================================================== the arena
class LoraArena:
    """Flat fp32 storage for every LoRA parameter of a model, its gradient and the AdamW moments.

    Layout = the reference's flat `unet_lora.pt` order `[up_0, down_0, up_1, down_1, ...]` (utils/lora.py:581-594), each
    tensor contiguous, the total padded to a multiple of 4 values.  `params`, `grads`, `exp_avg`, `exp_avg_sq` are four
    buffers of identical layout; the layers' `lora_up.weight` / `lora_down.weight` are views into `params` and their
    `.grad` views into `grads`, so a torch optimizer sees ordinary parameters while the kernels (and NCCL) see one buffer.
    """

    def __init__(self, shapes, device):
        self.shapes = [tuple(s) for s in shapes]
        self.offsets, off = [], 0
        for s in self.shapes:
            self.offsets.append(off)
            off += math.prod(s)
        self.numel = off
        self.padded = (off + 3) // 4 * 4
        self.params = torch.zeros(self.padded, device=device, dtype=torch.float32)
        self.grads = torch.zeros(self.padded, device=device, dtype=torch.float32)
        self.exp_avg = torch.zeros(self.padded, device=device, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(self.padded, device=device, dtype=torch.float32)
        self.step = 0

    def view(self, buf, i):
        s = self.shapes[i]
        return buf[self.offsets[i]:self.offsets[i] + math.prod(s)].view(s)

    def param(self, i):
        return self.view(self.params, i)

    def grad(self, i):
        return self.view(self.grads, i)

    def zero_grad(self):
        self.grads.zero_()

    def grad_norm(self, grad_scale=1.0):
        """L2 norm of the (scaled) gradient arena: one reduction kernel (accelerator.clip_grad_norm_, :1191)."""
        return (ops.sum_squares(self.grads).sqrt() * abs(grad_scale))

    def adamw_step(self, *, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, grad_scale=1.0, max_grad_norm=None):
        """One fused AdamW launch over the whole arena (torch.optim.AdamW semantics).  grad_scale folds the DDP mean
        (1 / world) in; max_grad_norm clips by the global norm like clip_grad_norm_ (the clip factor is computed on the
        device and read back: one 4-byte sync, as the reference's clip does)."""
        if max_grad_norm is not None:
            total = float(self.grad_norm(grad_scale))
            grad_scale = grad_scale * min(1.0, max_grad_norm / (total + 1e-6))
        self.step += 1
        ops.adamw_step(self.params, self.grads, self.exp_avg, self.exp_avg_sq, lr=lr, betas=betas, eps=eps,
                       weight_decay=weight_decay, step=self.step, grad_scale=grad_scale)

    # wire format (utils/lora.py:581-594): the flat list of tensors
    def to_list(self):
        return [self.param(i).detach().clone().cpu() for i in range(len(self.shapes))]

    @torch.no_grad()
    def load_list(self, tensors):
        assert len(tensors) == len(self.shapes)
        for i, t in enumerate(tensors):
            assert tuple(t.shape) == self.shapes[i], (i, tuple(t.shape), self.shapes[i])
            self.param(i).copy_(t)


# =============================================================================== functional forward / backward
def _keep_mask(shape, p, device, generator=None):
    if p <= 0.0:
        return None, 1.0
    keep = torch.empty(shape, device=device, dtype=torch.uint8).bernoulli_(1.0 - p, generator=generator)   # 1 byte / element
    return keep, 1.0 / (1.0 - p)


class _PackedLora:
    """bf16 GEMM operands of one layer, re-derived from the fp32 arena views after every optimizer step."""

    def __init__(self, kind, w, bias, up, down, scale):
        self.kind, self.scale = kind, float(scale)
        self.up_f32, self.down_f32 = up, down
        wd = w.detach()
        self.bias = bias.detach().float().contiguous() if bias is not None else None
        self.cout, self.cin = wd.shape[0], wd.shape[1]
        self.r = up.shape[1]
        if self.cin % 64 or self.cout % 64 or self.r % 64:
            raise NotImplementedError(f"LoRA layer {self.cin} -> {self.cout} (rank {self.r}): the tensor-core path needs channel "
                                      "counts and rank in multiples of 64 (every VC2 layer except the 4-channel conv_in / out)")
        # dgrad: dx = dy W + dt D is ONE implicit GEMM over the channel-concatenated pair (dy | dt) — the A operand's second
        # source — against wd_t = [W^T | D^T] concatenated along K (per tap for the convolutions): no second pass over dx.
        if kind == "linear":   # nn.Linear, or a 1x1 convolution (ResBlock skip_connection) run as a per-pixel Linear
            wd = wd.reshape(wd.shape[0], -1)
            self.w = wd.to(BF16).contiguous()                                   # [N, K]
            self.taps = 1
            self.wd_t = torch.empty((self.cin, self.cout + self.r), device=wd.device, dtype=BF16)     # [K, N + r]
            self.wd_t[:, :self.cout] = wd.t()
        else:
            sp = tuple(range(2, wd.dim()))
            self.w = ops.pack_conv_weight(wd)                                   # [Cout, taps * Cin]
            self.taps = math.prod(wd.shape[2:])
            w_t = ops.pack_conv_weight(wd.transpose(0, 1).flip(sp).contiguous())   # [Cin, taps * Cout], taps reversed
            self.wd_t = torch.empty((self.cin, self.taps * (self.cout + self.r)), device=wd.device, dtype=BF16)
            self.wd_t.view(self.cin, self.taps, self.cout + self.r)[:, :, :self.cout] = w_t.view(self.cin, self.taps, self.cout)
        self.refresh()

    @torch.no_grad()
    def refresh(self):
        """bf16 GEMM operands from the fp32 LoRA weights.  After the first call the operands are rewritten IN PLACE: CUDA
        graphs captured over this layer keep reading the same buffers."""
        up, down = self.up_f32, self.down_f32
        up2 = up.reshape(self.cout, self.r)
        if self.kind == "linear":
            down2 = down.reshape(self.r, -1)
            vals = dict(u=up2, u_t=up2.t(), d=down2, d_t=down2.t())
        else:
            sp = tuple(range(2, down.dim()))
            vals = dict(u=up2, u_t=up2.t(), d=ops.pack_conv_weight(down),                          # [r, taps * Cin]
                        d_t=ops.pack_conv_weight(down.transpose(0, 1).flip(sp).contiguous()))      # [Cin, taps * r], taps reversed
        d_t = vals.pop("d_t")
        self.wd_t.view(self.cin, self.taps, self.cout + self.r)[:, :, self.cout:] = d_t.reshape(self.cin, self.taps, self.r)
        for name, v in vals.items():
            cur = getattr(self, name, None)
            if cur is None:
                setattr(self, name, v.to(BF16).contiguous())
            else:
                cur.copy_(v)


def _base_op(kind, x, w, bias, residual=None, bias_div=None):
    """The forward implicit GEMM of the three layer kinds on channels-last bf16 (x: [M,K] | [n,h,w,C] | [b,t,hw,C]).
    conv2d: a 2-D bias [rows, Cout] with bias_div carries a per-sample bias row (frame // bias_div): the ResBlock's
    conv bias + timestep-embedding row (openaimodel3d.py:237-246)."""
    if kind == "linear":
        return ops.linear(x, w, bias, residual=residual)
    if kind == "conv2d":
        if bias is not None and bias.dim() == 1:
            bias, bias_div = bias.view(1, -1), x.shape[0]
        return ops.conv3x3(x, w, bias, bias_div=bias_div or 1, residual=residual)
    return ops.tconv3(x, w, bias, residual=residual)


def lora_forward(pk: _PackedLora, x, mask, mask_scale, bias_rows=None, bias_div=None, drop_p=0.0, addend=None):
    """-> (y, t, mask, mask_scale): t = lora_down(x) is kept for the backward.  Either the caller supplies the keep-mask
    (mask, mask_scale) or — drop_p > 0 — it is drawn inside the scale kernel (`ops.dropout_scale`: no mask kernel, no mask
    read in the forward) and returned for the backward.  bias_rows / bias_div: see _base_op (conv2d only).  addend ([M, Cout] bf16):
    the residual the caller would add to y next (`x + attn(...)`, `h + out_layers(...)`); y then already contains it — folded
    into the dropout pass when there is one, a separate add otherwise."""
    t = _base_op(pk.kind, x, pk.d, None)                                         # [.., r]
    u = ops.linear(t.view(-1, pk.r), pk.u, None)                                 # [M, Cout]
    if drop_p > 0.0 and mask is None:
        branch, mask = ops.dropout_scale(u, drop_p, pk.scale, addend=addend)
        mask_scale = 1.0 / (1.0 - drop_p)
        addend = None
    else:
        branch = ops.scale_mask(u, pk.scale * mask_scale, mask)
    bias = pk.bias if bias_rows is None else bias_rows
    y = _base_op(pk.kind, x, pk.w, bias, residual=branch.view(*x.shape[:-1], pk.cout), bias_div=bias_div)
    if addend is not None:
        y = ops.add(y.view(-1, pk.cout), addend.view(-1, pk.cout)).view(y.shape)
    return y, t, mask, mask_scale


def lora_backward(pk: _PackedLora, x, t, mask, mask_scale, dy, g_up, g_down, need_dx=True):
    """dy: channels-last bf16 like y.  Accumulates dU / dD into the fp32 arena views g_up / g_down; returns dx."""
    dy = dy.contiguous()
    du = ops.scale_mask(dy.view(-1, pk.cout), pk.scale * mask_scale, mask)       # [M, Cout]
    t2 = t.view(-1, pk.r)
    # lora_up.weight [Cout, r(,1,1..)]: element (c = cout, j) at c * r + j
    ops.wgrad(du, t2, g_up, out_strides=(1, pk.r, 0))
    dt = ops.linear(du, pk.u_t, None).view(*x.shape[:-1], pk.r)                  # [.., r]
    # lora_down.weight [r, Cin, taps]: element (j, c, tap) at j * Cin * taps + c * taps + tap
    taps = None if pk.kind == "linear" else (_TAPS_3X3 if pk.kind == "conv2d" else _TAPS_T3)
    n_taps = 1 if taps is None else len(taps)
    ops.wgrad(x, dt, g_down, taps=taps, out_strides=(pk.cin * n_taps, n_taps, 1))
    if not need_dx:
        return None
    return _base_op(pk.kind, (dy.view(*x.shape[:-1], pk.cout), dt), pk.wd_t, None)


class _LoraFn(torch.autograd.Function):
    """autograd glue: x in the layer's channels-last bf16 layout; the LoRA weights are passed so that autograd routes
    `.grad` to them, but their gradients are written by the kernels straight into the arena (returned as None here)."""

    @staticmethod
    def forward(ctx, x, layer, up, down):
        pk = layer._packed()
        p = layer.dropout_p if layer.training else 0.0
        y, t, mask, ms = lora_forward(pk, x, None, 1.0, drop_p=p)
        ctx.layer, ctx.mask, ctx.ms = layer, mask, ms
        ctx.save_for_backward(x, t)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, t = ctx.saved_tensors
        layer = ctx.layer
        layer._own_grads()
        dx = lora_backward(layer._packed(), x, t, ctx.mask, ctx.ms, dy.to(BF16), layer.lora_up.weight.grad, layer.lora_down.weight.grad,
                           need_dx=ctx.needs_input_grad[0])
        return dx, None, None, None


# =============================================================================== modules (reference names and attributes)
class _LoraBase(nn.Module):
    kind = "linear"

    def _finish_init(self, base, r, dropout_p, scale, down_shape, up_shape):
        self.r, self.dropout_p, self.scale = r, dropout_p, scale
        self.dropout = nn.Dropout(dropout_p)
        self.selector = nn.Identity()
        self._pk = None
        for p in base.parameters():
            p.requires_grad_(False)
        nn.init.normal_(self.lora_down.weight, std=1 / r)
        nn.init.zeros_(self.lora_up.weight)

    def base_layer(self):
        return self.linear if self.kind == "linear" else self.conv

    def bind_arena(self, arena: LoraArena, i_up: int, i_down: int):
        """Move the LoRA weights (and their .grad) into the arena: views, no copies afterwards."""
        with torch.no_grad():
            arena.param(i_up).copy_(self.lora_up.weight)
            arena.param(i_down).copy_(self.lora_down.weight)
        self.lora_up.weight = nn.Parameter(arena.param(i_up))
        self.lora_down.weight = nn.Parameter(arena.param(i_down))
        self.lora_up.weight.grad = arena.grad(i_up)
        self.lora_down.weight.grad = arena.grad(i_down)
        self._pk = None

    def _own_grads(self):
        for prm in (self.lora_up.weight, self.lora_down.weight):
            if prm.grad is None:
                prm.grad = torch.zeros_like(prm)

    def _packed(self):
        if self._pk is None:
            base = self.base_layer()
            if base.weight.device.type != "cuda":
                raise RuntimeError("LoraInjected*(B200) runs on a CUDA device only (no CPU fallback)")
            self._own_grads()
            self._pk = _PackedLora(self.kind, base.weight, base.bias, self.lora_up.weight.detach(), self.lora_down.weight.detach(),
                                   self.scale)
        return self._pk

    def refresh(self):
        """Re-derive the bf16 operands after the fp32 LoRA weights changed (optimizer step / load)."""
        if self._pk is not None:
            self._pk.refresh()

    def _apply(self, fn, *a, **k):
        self._pk = None
        return super()._apply(fn, *a, **k)

    def realize_as_lora(self):
        return self.lora_up.weight.data * self.scale, self.lora_down.weight.data


class LoraInjectedLinear(_LoraBase):
    kind = "linear"

    def __init__(self, in_features, out_features, bias=False, r=4, dropout_p=0.1, scale=1.0):
        super().__init__()
        r = min(r, in_features, out_features)
        self.linear = nn.Linear(in_features, out_features, bias)
        self.lora_down = nn.Linear(in_features, r, bias=False)
        self.lora_up = nn.Linear(r, out_features, bias=False)
        self._finish_init(self.linear, r, dropout_p, scale, None, None)

    def forward(self, input):
        if not input.is_cuda:
            raise RuntimeError("LoraInjectedLinear(B200): input must be a CUDA tensor (no CPU fallback)")
        self._packed()
        x = input.reshape(-1, input.shape[-1]).to(BF16).contiguous()
        y = _LoraFn.apply(x, self, self.lora_up.weight, self.lora_down.weight)
        return y.view(*input.shape[:-1], y.shape[-1]).to(input.dtype)


class LoraInjectedConv2d(_LoraBase):
    kind = "conv2d"

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True, r=4,
                 dropout_p=0.1, scale=1.0):
        super().__init__()
        ks = (kernel_size, kernel_size) if isinstance(kernel_size, int) else tuple(kernel_size)
        if ks != (3, 3) or stride not in (1, (1, 1)) or padding not in (1, (1, 1)) or dilation not in (1, (1, 1)) or groups != 1:
            raise NotImplementedError("LoraInjectedConv2d(B200): 3x3 / stride 1 / padding 1 convolutions (the ResBlock convs)")
        r = min(r, in_channels, out_channels)
        self.conv = nn.Conv2d(in_channels, out_channels, 3, padding=1, bias=bias)
        self.lora_down = nn.Conv2d(in_channels, r, 3, padding=1, bias=False)
        self.lora_up = nn.Conv2d(r, out_channels, 1, bias=False)
        self._finish_init(self.conv, r, dropout_p, scale, None, None)

    def forward(self, input):   # NCHW like the reference
        if not input.is_cuda:
            raise RuntimeError("LoraInjectedConv2d(B200): input must be a CUDA tensor (no CPU fallback)")
        self._packed()
        x = input.permute(0, 2, 3, 1).to(BF16).contiguous()
        y = _LoraFn.apply(x, self, self.lora_up.weight, self.lora_down.weight)
        return y.permute(0, 3, 1, 2).to(input.dtype)


class LoraInjectedConv3d(_LoraBase):
    kind = "conv3d"

    def __init__(self, in_channels, out_channels, kernel_size=(3, 1, 1), padding=(1, 0, 0), bias=False, r=4, dropout_p=0.0, scale=1.0):
        super().__init__()
        if tuple(kernel_size) != (3, 1, 1) or tuple(padding) != (1, 0, 0):
            raise NotImplementedError("LoraInjectedConv3d(B200): (3,1,1) / padding (1,0,0) convolutions (TemporalConvBlock)")
        r = min(r, in_channels, out_channels)
        self.conv = nn.Conv3d(in_channels, out_channels, (3, 1, 1), padding=(1, 0, 0), bias=bias)
        self.lora_down = nn.Conv3d(in_channels, r, (3, 1, 1), padding=(1, 0, 0), bias=False)
        self.lora_up = nn.Conv3d(r, out_channels, 1, bias=False)
        self._finish_init(self.conv, r, dropout_p, scale, None, None)

    def forward(self, input):   # [B, C, T, H, W] like the reference
        if not input.is_cuda:
            raise RuntimeError("LoraInjectedConv3d(B200): input must be a CUDA tensor (no CPU fallback)")
        self._packed()
        b, c, t, h, w = input.shape
        x = input.permute(0, 2, 3, 4, 1).reshape(b, t, h * w, c).to(BF16).contiguous()
        y = _LoraFn.apply(x, self, self.lora_up.weight, self.lora_down.weight)
        return y.view(b, t, h, w, -1).permute(0, 4, 1, 2, 3).to(input.dtype)


# =============================================================================== whole-model census / arena construction
def arena_for_unet(unet: nn.Module, device, r: int = 64) -> LoraArena:
    """The gradient arena of a UNet's 575 LoRA target layers in the reference's flat-list order (lora.lora_shapes):
    117 142 176 fp32 values for the VC2 config at r = 64."""
    from .lora import lora_shapes
    shapes = []
    for up, down in lora_shapes(unet, r):
        shapes += [up, down]
    return LoraArena(shapes, device)


def unet_lora_workload(unet: nn.Module, batch=1, frames=16, h=40, w=64, ctx_len=77):
    """The LoRA target layers of a (B200) UNetModel with the activation geometry each one sees in a forward on
    [batch, 4, frames, h, w] latents: [(name, kind, points, cin, cout)] in `named_modules()` order (= arena order), where
    kind is linear | conv2d | conv3d | skip (not on the tensor-core training path: the 4-channel conv_in / out and the
    strided / upsampling convolutions) and points is the channels-last point grid ([M] | [n, h, w] | [b, t, hw])."""
    from .lora import lora_target_layers
    from . import unet as U
    geo = {}
    n = batch * frames

    def walk(seq, hh, ww):
        for layer in seq:
            if isinstance(layer, (U.ResBlock, U.SpatialTransformer, U.TemporalTransformer, U.Downsample, U.Upsample)):
                geo[layer] = (hh, ww)
            if isinstance(layer, U.Downsample):
                hh, ww = hh // 2, ww // 2
            elif isinstance(layer, U.Upsample):
                hh, ww = hh * 2, ww * 2
        return hh, ww
    hh, ww = h, w
    for blk in unet.input_blocks:
        hh, ww = walk(blk, hh, ww)
    if getattr(unet, "init_attn", None) is not None:
        walk(unet.init_attn, h, w)
    hh, ww = walk(unet.middle_block, hh, ww)
    for blk in unet.output_blocks:
        hh, ww = walk(blk, hh, ww)
    owner = {}
    for mod, g in geo.items():
        for sub in mod.modules():
            owner[sub] = (mod, g)
    out = []
    for name, m in lora_target_layers(unet):
        cin = m.in_features if isinstance(m, nn.Linear) else m.in_channels
        cout = m.out_features if isinstance(m, nn.Linear) else m.out_channels
        if m not in owner:                       # time_embed / fps_embedding / cond projections: one row per sample
            kind, pts = ("linear", (batch,)) if isinstance(m, nn.Linear) else ("skip", None)
        else:
            mod, (gh, gw) = owner[m]
            if isinstance(mod, (U.Downsample, U.Upsample)):
                kind, pts = "skip", None
            elif isinstance(m, nn.Conv3d):
                kind, pts = "conv3d", (batch, frames, gh * gw)
            elif isinstance(m, nn.Conv2d):
                kind, pts = ("conv2d", (n, gh, gw)) if m.kernel_size == (3, 3) else ("linear", (n * gh * gw,))
            elif ".emb_layers." in name:
                kind, pts = "linear", (batch,)
            elif name.endswith(("attn2.to_k", "attn2.to_v")) and isinstance(mod, U.SpatialTransformer):
                kind, pts = "linear", (n * ctx_len,)   # the reference projects the frame-repeated text context (openaimodel3d.py:710)
            else:
                kind, pts = "linear", (n * gh * gw,)
        if kind != "skip" and (cin % 64 or cout % 64):
            kind, pts = "skip", None
        out.append((name, kind, pts, cin, cout))
    return out
