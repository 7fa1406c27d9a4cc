"""The student UNet of the v2 FULL fine-tune step on B200 (train_latent_t2v_turbo_v2.py:799-840 parameter groups, :1049-1056
student forward, :1239-1247 target forward, :1264-1276 backward / clip / AdamW / EMA): every parameter of the VideoCrafter2
UNet trains — weights, biases and norm affines, 1.41 B values — and `motion_cond` feeds the embedding (openaimodel3d.py:690-697).

`FullUNet(unet)` is the training view of a (B200) `UNetModel`, sharing `train_unet.StudentUNet`'s traversal (the same forward
arithmetic, tapes and input-gradient kernels) with a different layer type and the affine gradients added:

    GEMM layers     y = W x (+ b)                   one tcgen05 implicit GEMM (the base op of the LoRA layer, no side branch)
                    dx = dy W                       the same GEMM on the transposed / tap-reversed weight
                    dW += dy^T x (per tap)          t2v_wgrad over 64-column slices of dy read in place (ops.wgrad_wide)
                    db += colsum(dy)                t2v_colsum_samples
    GroupNorm       dx: t2v_groupnorm_bwd           dgamma / dbeta: t2v_groupnorm_affine_grad (statistics from dx's workspace)
    LayerNorm       dx: t2v_layernorm_bwd           dgamma / dbeta: t2v_layernorm_affine_grad

All parameters, gradients and AdamW moments live in ONE flat fp32 arena each (`FullArena`), laid out in the order in which the
backward completes blocks (reversed), so that the data-parallel exchange — a bucketed all-reduce over 5.65 GB, dist.ArenaReducer —
starts on the output blocks' gradients while the backward is still in the encoder; the two optimizer groups of the reference
(:799-840, other / temporal: lr and lr * temporal_lr_scale) interleave along that order and are stepped as one fused AdamW launch
per contiguous run (33 on the VC2 UNet), gradient clipping is one reduction and the EMA target (`update_ema`,
utils/common_utils.py:308-319) one elementwise launch.  The `UNetModel`'s nn.Parameters are re-pointed at
the arena (views), so `unet.state_dict()` is the `unet.pt` wire format at any time; the bf16 GEMM operands are re-derived from
the fp32 arena after every optimizer step (`refresh`).

Status: the whole composition is checked on CPU against the UNMODIFIED reference's autograd through a restatement
of every kernel contract (tests/mock_ops.py); on B200 the three new kernels (train_full.cu) and `wgrad_wide` pass their contract
tests and one whole step agrees with the reference composition; the step has not been timed and its backward has not yet been
pinned under a linear loss on the device (tests/test_zz_full_train_gpu.py).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import ops
from . import unet as U
from .lora_train import _base_op
from .train_unet import StudentUNet, _Norm, _require_cuda

BF16 = torch.bfloat16
_TAPS_3X3 = ops._TAPS_3X3
_TAPS_T3 = ops._TAPS_T3


# =============================================================================== parameter groups + arena
def param_groups(unet: nn.Module):
    """-> (other_names, temporal_names): the reference's two optimizer groups, by ITS rule (train_latent_t2v_turbo_v2.py:799-815):
    `init_attn.0.*` is temporal; otherwise a parameter is temporal iff the module named by the first THREE components of its
    name is a TemporalTransformer (so the temporal transformers of input / output blocks are, the one inside middle_block —
    `middle_block.2`, two components deep — is not: its three-component prefix names a child module)."""
    mods = dict(unet.named_modules())
    other, temporal = [], []
    for n, _ in unet.named_parameters():
        parts = n.split(".")
        if n.startswith("init_attn.0"):
            temporal.append(n)
        elif len(parts) > 2 and isinstance(mods.get(".".join(parts[:3])), U.TemporalTransformer):
            temporal.append(n)
        else:
            other.append(n)
    return other, temporal


def _depth_key(name: str):
    """Arena position of a parameter: the order in which the BACKWARD finishes blocks, reversed — embeddings, input blocks,
    init_attn, middle block, output blocks, out — so that gradients become final from the END of the arena towards its start
    (what dist.ArenaReducer's bucketed all-reduce overlaps with; the same convention as the LoRA arena of the v1 step)."""
    parts = name.split(".")
    top = parts[0]
    rank = {"input_blocks": 1, "init_attn": 2, "middle_block": 3, "output_blocks": 4, "out": 5}.get(top, 0)
    idx = int(parts[1]) if top in ("input_blocks", "output_blocks") else 0
    return (rank, idx)


class FullArena:
    """Flat fp32 storage of every UNet parameter, its gradient and the AdamW moments (+ optionally the EMA target).  Tensors are
    laid out in backward-completion order (`_depth_key`; stable within a block), every tensor starts on a 16-byte boundary
    (kernels read gamma / beta / bias rows vectorised).  The reference's two optimizer groups (other / temporal, `param_groups`)
    interleave along the arena: `runs` lists the maximal contiguous ranges of one group, each one fused AdamW launch."""

    def __init__(self, unet: nn.Module, device, with_target=False):
        other, temporal = param_groups(unet)
        params = dict(unet.named_parameters())
        tset = set(temporal)
        order = {n: i for i, n in enumerate(params)}
        self.names = sorted(params, key=lambda n: (_depth_key(n), order[n]))
        self.index = {n: i for i, n in enumerate(self.names)}
        self.is_temporal = [n in tset for n in self.names]
        self.shapes, self.offsets, off = [], [], 0
        for n in self.names:
            self.shapes.append(tuple(params[n].shape))
            self.offsets.append(off)
            off += (params[n].numel() + 3) // 4 * 4
        self.numel = sum(math.prod(s) for s in self.shapes)
        self.padded = off
        self.runs = []                      # [(lo, hi, temporal?)]
        for i, t in enumerate(self.is_temporal):
            hi = self.offsets[i + 1] if i + 1 < len(self.names) else off
            if self.runs and self.runs[-1][2] == t:
                self.runs[-1] = (self.runs[-1][0], hi, t)
            else:
                self.runs.append((self.offsets[i], hi, t))
        self.params = torch.zeros(off, device=device, dtype=torch.float32)
        self.grads = torch.zeros(off, device=device, dtype=torch.float32)
        self.exp_avg = torch.zeros(off, device=device, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(off, device=device, dtype=torch.float32)
        self.step = 0
        with torch.no_grad():
            for i, n in enumerate(self.names):
                self.view(self.params, i).copy_(params[n].detach().to(device=device, dtype=torch.float32))
        self.target = self.params.clone() if with_target else None

    def view(self, buf, i):
        s = self.shapes[i]
        return buf[self.offsets[i]:self.offsets[i] + math.prod(s)].view(s)

    def param(self, name):
        return self.view(self.params, self.index[name])

    def grad(self, name):
        return self.view(self.grads, self.index[name])

    def first_offset(self, prefix: str) -> int:
        """Offset of the first tensor whose name starts with `prefix + "."` (a top-level block of the UNet)."""
        for n, o in zip(self.names, self.offsets):
            if n.startswith(prefix + "."):
                return o
        raise KeyError(prefix)

    def bind(self, unet: nn.Module, buf=None):
        """Re-point the module's nn.Parameters at this arena (views of `buf`, default the live parameters): state_dict(),
        load_state_dict() and the inference forward of that module then read / write the arena directly."""
        buf = self.params if buf is None else buf
        for n, p in unet.named_parameters():
            p.data = self.view(buf, self.index[n])
            p.requires_grad_(False)

    def zero_grad(self):
        self.grads.zero_()

    def grad_norm(self, grad_scale=1.0):
        return ops.sum_squares(self.grads).sqrt() * abs(grad_scale)

    def adamw_step(self, *, lr, temporal_lr_scale=1.0, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, grad_scale=1.0, max_grad_norm=None):
        """torch.optim.AdamW over the two groups of :829-840 (lr, lr * temporal_lr_scale), one fused launch per contiguous run of a
        group; clip_grad_norm_ over ALL parameters first (:1267), the clip factor and the data-parallel mean folded into grad_scale.
        One deliberate difference from torch: a parameter that received NO gradient in a step (torch: `.grad is None`, skipped) is
        stepped here with a zero gradient — its moments decay and, if weight_decay > 0, it is decayed.  That only concerns
        motion_cond_proj / combine_proj when a motion-conditioned model is trained with use_motion_cond off; the v2 script's default
        weight decay is 0."""
        if max_grad_norm is not None:
            total = float(self.grad_norm(grad_scale))
            grad_scale = grad_scale * min(1.0, max_grad_norm / (total + 1e-6))
        self.step += 1
        for lo, hi, temporal in self.runs:
            ops.adamw_step(self.params[lo:hi], self.grads[lo:hi], self.exp_avg[lo:hi], self.exp_avg_sq[lo:hi],
                           lr=lr * temporal_lr_scale if temporal else lr, betas=betas, eps=eps, weight_decay=weight_decay, step=self.step,
                           grad_scale=grad_scale)

    # ---- resume (train_latent_t2v_turbo_v2.py:1283-1313 saves unet.pt / the accelerator state every `checkpointing_steps`)
    def state_dict(self):
        """Optimizer-side state for resuming: AdamW moments, step count and the EMA target, keyed by PARAMETER NAME so that it is
        independent of the arena's internal order (the parameters themselves are `unet.state_dict()`, the `unet.pt` wire format)."""
        out = {"step": self.step, "exp_avg": {}, "exp_avg_sq": {}}
        if self.target is not None:
            out["target"] = {}
        for i, n in enumerate(self.names):
            out["exp_avg"][n] = self.view(self.exp_avg, i).detach().cpu().clone()
            out["exp_avg_sq"][n] = self.view(self.exp_avg_sq, i).detach().cpu().clone()
            if self.target is not None:
                out["target"][n] = self.view(self.target, i).detach().cpu().clone()
        return out

    @torch.no_grad()
    def load_state_dict(self, sd):
        if set(sd["exp_avg"]) != set(self.names) or set(sd["exp_avg_sq"]) != set(self.names):
            raise KeyError("FullArena.load_state_dict: parameter names differ from this model's")
        self.step = int(sd["step"])
        for i, n in enumerate(self.names):
            self.view(self.exp_avg, i).copy_(sd["exp_avg"][n])
            self.view(self.exp_avg_sq, i).copy_(sd["exp_avg_sq"][n])
            if self.target is not None and "target" in sd:
                self.view(self.target, i).copy_(sd["target"][n])

    def ema_step(self, decay):
        """update_ema(target_unet.parameters(), unet.parameters(), decay) (:1273-1276) over the whole arena in one launch."""
        if self.target is None:
            raise RuntimeError("FullArena was built without an EMA target (with_target=True)")
        ops.ema_update(self.target, self.params, decay)


# =============================================================================== one trainable GEMM layer
class _FullLayer:
    """A GEMM layer whose own weight (and bias) train.  kind: linear | conv2d | conv3d (nn.Conv1d k = 1 and 1x1 Conv2d run as
    per-point Linear layers).  pad_in / pad_out: the 4-channel conv_in / out run zero-padded to 64 channels; the padding rows
    of their gradients are dropped when they are folded into the arena."""
    lora = False

    def __init__(self, name, module, kind, arena: FullArena, pad_in=0, pad_out=0):
        self.name, self.module, self.kind = name, module, kind
        self.arena = arena
        self.w32 = arena.param(name + ".weight")
        self.gw = arena.grad(name + ".weight")
        has_b = module.bias is not None
        self.b32 = arena.param(name + ".bias") if has_b else None
        self.gb = arena.grad(name + ".bias") if has_b else None
        self.cout, self.cin = self.w32.shape[0], self.w32.shape[1]
        self.taps = math.prod(self.w32.shape[2:]) if kind != "linear" else 1
        self.pad_in, self.pad_out = pad_in, pad_out
        self.pc_out, self.pc_in = pad_out or self.cout, pad_in or self.cin
        if self.pc_out % 8 or self.pc_in % 8:
            raise NotImplementedError(f"{name}: channel counts {self.cin} -> {self.cout} need padding to a multiple of 8")
        self.pk = self           # the ResBlock code reads `layer.pk.bias` (conv bias + embedding row, train_unet._res_fwd)
        self.w = self.w_t = self.bias = None

    @property
    def padded(self):
        return bool(self.pad_in or self.pad_out)

    def _w_padded(self):
        w = self.w32
        if self.kind == "linear":
            w = w.reshape(self.cout, -1)
        if not self.padded:
            return w
        out = torch.zeros((self.pc_out, self.pc_in) + tuple(w.shape[2:]), device=w.device, dtype=torch.float32)
        out[:self.cout, :self.cin] = w
        return out

    def pack(self):
        dev = self.w32.device
        k = self.taps * self.pc_in
        self.w = torch.empty((self.pc_out, k), device=dev, dtype=BF16)                    # forward operand [N, taps * K]
        self.w_t = torch.empty((self.pc_in, self.taps * self.pc_out), device=dev, dtype=BF16)   # dgrad operand (taps reversed)
        if self.padded:
            self._gw_p = torch.zeros((self.pc_out, self.pc_in) + tuple(self.w32.shape[2:]), device=dev, dtype=torch.float32)
            self._gb_p = torch.zeros(self.pc_out, device=dev, dtype=torch.float32) if self.b32 is not None else None
            self.bias = torch.zeros(self.pc_out, device=dev, dtype=torch.float32) if self.b32 is not None else None
        else:
            self.bias = self.b32          # the arena view itself: the bias needs no refresh
        self.refresh()

    @torch.no_grad()
    def refresh(self):
        """bf16 GEMM operands from the fp32 arena (in place: captured CUDA graphs keep reading the same buffers)."""
        w = self._w_padded()
        if self.kind == "linear":
            self.w.copy_(w)
            self.w_t.copy_(w.t())
        else:
            sp = tuple(range(2, w.dim()))
            self.w.copy_(ops.pack_conv_weight(w))
            self.w_t.copy_(ops.pack_conv_weight(w.transpose(0, 1).flip(sp).contiguous()))
        if self.padded and self.b32 is not None:
            self.bias[:self.cout].copy_(self.b32)

    # x: channels-last bf16 in the layer's point grid ([M, K] | [n, h, w, C] | [b, t, hw, C])
    def forward(self, x, training, bias_rows=None, bias_div=None, addend=None):
        bias = self.bias if bias_rows is None else bias_rows
        res = addend.view(*x.shape[:-1], self.pc_out) if addend is not None else None
        return _base_op(self.kind, x, self.w, bias, residual=res, bias_div=bias_div), (x,)

    def backward(self, saved, dy, need_dx=True):
        (x,) = saved
        dy = dy.contiguous().view(*x.shape[:-1], self.pc_out)
        gw = self._gw_p if self.padded else self.gw
        taps = None if self.kind == "linear" else (_TAPS_3X3 if self.kind == "conv2d" else _TAPS_T3)
        # weight [Cout, Cin, taps]: element (j = cout, c = cin, tap) at j * Cin * taps + c * taps + tap
        ops.wgrad_wide(x, dy, gw, taps=taps, out_strides=(self.pc_in * self.taps, self.taps, 1))
        if self.b32 is not None:
            gb = self._gb_p if self.padded else self.gb
            ops.colsum_samples(dy.view(-1, self.pc_out), dy.numel() // self.pc_out, out=gb.view(1, -1))
        if self.padded:
            self.gw.add_(self._gw_p[:self.cout, :self.cin].reshape(self.gw.shape))
            self._gw_p.zero_()
            if self.b32 is not None:
                self.gb.add_(self._gb_p[:self.cout])
                self._gb_p.zero_()
        if not need_dx:
            return None
        return _base_op(self.kind, dy, self.w_t, None)


class _TrainNorm(_Norm):
    """GroupNorm / LayerNorm whose affine parameters train: w / b are the arena's fp32 views, gw / gb their gradients."""

    def __init__(self, name, m, arena: FullArena):
        self.module = m
        self.eps = m.eps
        self.w, self.b = arena.param(name + ".weight"), arena.param(name + ".bias")
        self.gw, self.gb = arena.grad(name + ".weight"), arena.grad(name + ".bias")


# =============================================================================== the v2 student
class FullUNet(StudentUNet):
    def __init__(self, unet: U.UNetModel, tconv_dropout: float = 0.1, with_target: bool = False):
        dev = unet.time_embed[0].weight.device
        _require_cuda("FullUNet", dev)
        self.unet, self.device = unet, dev
        self.training = True
        self.tconv_p = tconv_dropout
        self.arena = FullArena(unet, dev, with_target=with_target)
        self.arena.bind(unet)
        self._mod_name = {m: n for n, m in unet.named_modules()}
        self.layers, self.layer_list, self._plain = {}, [], {}
        self._packed = False
        self._build()
        self._train_norms()
        # first arena offset of every top-level block: once the backward has passed a block, every gradient at or above that
        # offset is final (the arena is in backward-completion order) -> dist.ArenaReducer.ready overlaps the 5.65 GB exchange
        self._first_offset = {"middle_block": self.arena.first_offset("middle_block")}
        for j in range(len(unet.output_blocks)):
            self._first_offset[f"output_blocks.{j}"] = self.arena.first_offset(f"output_blocks.{j}")
        self.on_grads_final = None      # callable(offset)

    # ------------------------------------------------------------------ structure
    def _L(self, m):
        lay = self.layers.get(m)
        if lay is None:
            name = self._mod_name[m]
            if isinstance(m, nn.Linear) or isinstance(m, nn.Conv1d) or (isinstance(m, nn.Conv2d) and m.kernel_size == (1, 1)):
                kind = "linear"
            elif isinstance(m, nn.Conv2d):
                kind = "conv2d"
            elif isinstance(m, nn.Conv3d):
                kind = "conv3d"
            else:
                raise TypeError(f"{name}: {type(m)}")
            cin = m.in_features if isinstance(m, nn.Linear) else m.in_channels
            cout = m.out_features if isinstance(m, nn.Linear) else m.out_channels
            lay = _FullLayer(name, m, kind, self.arena, pad_in=64 if cin < 64 else 0, pad_out=64 if cout < 64 else 0)
            self.layers[m] = lay
            self.layer_list.append(lay)
        return lay

    def _motion_struct(self, u):
        if u.motion_cond_proj is None:
            return None
        return (self._L(u.motion_cond_proj), self._L(u.combine_proj))

    def _train_norms(self):
        """Swap every frozen `_Norm` of the traversal structures for a `_TrainNorm` bound to the arena."""
        def tn(n):
            return _TrainNorm(self._mod_name[n.module], n.module, self.arena)

        def walk(items):
            for kind, S in items:
                if kind == "res":
                    S["gn1"], S["gn2"] = tn(S["gn1"]), tn(S["gn2"])
                    if S["tconv"] is not None:
                        S["tconv"] = [(tn(g), c) for g, c in S["tconv"]]
                elif kind in ("st", "tt"):
                    S["gn"] = tn(S["gn"])
                    S["ln"] = [tn(n) for n in S["ln"]]
        for seq in self.s_input + self.s_output + [self.s_middle] + ([self.s_init] if self.s_init is not None else []):
            walk(seq)
        self.s_out = (tn(self.s_out[0]), self.s_out[1])

    def pack(self):
        for lay in self.layer_list:
            lay.pack()
        half = self.unet.model_channels // 2
        self.freqs = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32) / half).to(self.device)
        self._packed = True

    def refresh(self):
        """Re-derive the bf16 GEMM operands after `arena.adamw_step` / a load into the arena."""
        if getattr(self, "_refresh_graph", None) is not None:
            self._refresh_graph.replay()
            return
        for lay in self.layer_list:
            lay.refresh()

    # ------------------------------------------------------------------ norm adjoints with affine gradients
    def _gn_bwd(self, norm, x, dy, *, rows_per_sample, silu, dx_add=None):
        ws = []
        dx = ops.groupnorm_bwd(x, dy, norm.w, norm.b, rows_per_sample=rows_per_sample, eps=norm.eps, silu=silu, dx_add=dx_add, keep_ws=ws)
        ops.groupnorm_affine_grad(x, dy, norm.w, norm.b, ws[0], norm.gw, norm.gb, rows_per_sample=rows_per_sample, eps=norm.eps, silu=silu)
        return dx

    def _ln_bwd(self, norm, x, dy, *, dx_add=None):
        dx = ops.layernorm_bwd(x, dy, norm.w, norm.eps, dx_add=dx_add)
        ops.layernorm_affine_grad(x, dy, norm.gw, norm.gb, norm.eps)
        return dx

    # ------------------------------------------------------------------ embeddings (with motion_cond, openaimodel3d.py:683-705)
    def _emb_fwd(self, timesteps, fps, timestep_cond, bsz, motion_cond=None):
        tr, dev = self.training, self.device
        ctx = {}
        t = timesteps.to(device=dev, dtype=torch.float32).reshape(-1)
        if t.numel() == 1 and bsz > 1:
            t = t.expand(bsz).contiguous()
        t_emb = ops.sinusoidal_embedding(t, self.freqs, round_bf16=True)
        if timestep_cond is not None:
            cond, ctx["cond"] = self.s_cond.forward(timestep_cond.to(device=dev, dtype=BF16).contiguous(), tr)
            if motion_cond is not None:
                if self.s_motion is None:
                    raise RuntimeError("motion_cond given but the UNet has no motion_cond_proj (motion_cond_proj_dim=None)")
                mproj, comb = self.s_motion
                mc, ctx["mproj"] = mproj.forward(motion_cond.to(device=dev, dtype=BF16).contiguous(), tr)
                cat = torch.cat([cond, mc], 1).contiguous()            # [B, 640]: plumbing
                cond, ctx["comb"] = comb.forward(cat, tr)
            t_emb = t_emb + cond.float()
        elif motion_cond is not None:
            raise AssertionError("motion_cond requires timestep_cond (openaimodel3d.py:691)")
        h0, ctx["t0"] = self.s_time[0].forward(t_emb.to(BF16).contiguous(), tr)
        a0 = ops.silu(h0)
        emb, ctx["t2"] = self.s_time[1].forward(a0, tr)
        ctx["h0"] = h0
        if self.s_fps is not None:
            if isinstance(fps, int) or (torch.is_tensor(fps) and fps.numel() == 1 and bsz > 1):
                fps_t = torch.full((bsz,), float(fps), device=dev, dtype=torch.float32)
            else:
                fps_t = fps.to(device=dev, dtype=torch.float32).reshape(-1)
            f_emb = ops.sinusoidal_embedding(fps_t, self.freqs, round_bf16=True).to(BF16).contiguous()
            f0, ctx["f0"] = self.s_fps[0].forward(f_emb, tr)
            f2, ctx["f2"] = self.s_fps[1].forward(ops.silu(f0), tr)
            ctx["fh0"] = f0
            emb = ops.add(emb, f2)
        ctx["emb"] = emb
        self._emb_ctx = ctx
        self._emb_act = ops.silu(emb)
        self._d_emb_act = None

    def _emb_bwd(self):
        ctx = self._emb_ctx
        if self._d_emb_act is None:
            return
        d_emb = ops.silu_bwd(ctx["emb"], self._d_emb_act)
        d_a0 = self.s_time[1].backward(ctx["t2"], d_emb)
        d_t = self.s_time[0].backward(ctx["t0"], ops.silu_bwd(ctx["h0"], d_a0), need_dx="cond" in ctx)
        if self.s_fps is not None:
            d_f = self.s_fps[1].backward(ctx["f2"], d_emb)
            self.s_fps[0].backward(ctx["f0"], ops.silu_bwd(ctx["fh0"], d_f), need_dx=False)
        if "cond" in ctx:
            d_cond = d_t                                                  # d(t_emb + cond) / d cond = 1
            if "comb" in ctx:
                mproj, comb = self.s_motion
                d_cat = comb.backward(ctx["comb"], d_cond)                # [B, 640]
                mc_ch = d_cat.shape[-1] // 2
                mproj.backward(ctx["mproj"], d_cat[:, mc_ch:].contiguous(), need_dx=False)
                d_cond = d_cat[:, :mc_ch].contiguous()
            self.s_cond.backward(ctx["cond"], d_cond, need_dx=False)
