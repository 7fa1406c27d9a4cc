"""The student UNet of the consistency-distillation step on B200: LoRA-injected forward AND backward of the whole
VideoCrafter2 UNet through the C ABI (train_t2v_turbo_v1_lora.py:640-656 injection, :1040-1048 student forward, :1190 backward).

`StudentUNet(unet)` is the training view of a (B200) `UNetModel`: the same parameters, every `nn.Linear / nn.Conv2d / nn.Conv3d`
below it wrapped by a LoRA pair in the reference's traversal order (`lora.lora_target_layers`, 575 layers; utils/lora.py:387-486),
all LoRA weights, gradients and AdamW moments in ONE flat fp32 arena (`lora_train.LoraArena`).

    forward(x, timesteps, context=..., fps=..., timestep_cond=...)  -> eps prediction [B, 4, T, H, W]
    backward(d_eps)                                                  -> LoRA gradients accumulated into arena.grads

The forward is the training-mode arithmetic of the reference modules (LoRA dropout and the TemporalConvBlock dropouts, their masks
drawn inside the scale kernel from a device-resident Philox seed; unfused
LayerNorm / GEGLU so that the backward has the tensors it needs), the backward is written out by hand — no autograd graph:

    GEMM layers     lora_train.lora_forward / lora_backward: 3 + 2 tcgen05 GEMMs, t2v_dropout_scale and 2 t2v_wgrad launches per layer
    GroupNorm+SiLU  t2v_groupnorm / t2v_groupnorm_bwd          LayerNorm   t2v_layernorm / t2v_layernorm_bwd
    attention       t2v_attn_fwd (+lse2) / t2v_attn_bwd        temporal    t2v_attn_short_fwd / t2v_attn_short_bwd
    GEGLU           t2v_geglu (forward and adjoint)            emb add     bias rows in the conv epilogue / t2v_colsum_samples
    Downsample      stride-1 LoRA conv + subsample / zero stuffing;  Upsample: nearest 2x + LoRA conv / 2x2 pooling adjoint
    conv_in / out   the 4-channel layers (LoRA rank 4) run zero-padded to 64 channels / rank 64 on the same tensor-core path;
                    the padding rows of their gradients are dropped when they are folded into the arena.

Only the LoRA weights receive gradients (`unet.requires_grad_(False)` + trainable LoRA, :640-656), so no base-weight, bias or
norm-affine gradient is ever formed.  Activations and gradients are channels-last bf16 (`[B*T, H, W, C]`), statistics fp32.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import ops
from . import unet as U
from .lora import lora_rank, lora_target_layers
from .lora_train import LoraArena, _PackedLora, lora_backward, lora_forward

BF16 = torch.bfloat16


def _f32(t):
    return t.detach().float().contiguous()


def _require_cuda(what, device):
    """The training views run on a CUDA device only: there is no CPU fallback of the product path."""
    if torch.device(device).type != "cuda":
        raise RuntimeError(f"{what}(B200) runs on a CUDA device only (no CPU fallback)")


# =============================================================================== one GEMM layer (LoRA-injected or plain)
class _Layer:
    """A GEMM layer of the training path.  kind: linear | conv2d | conv3d.  `lora` False = a frozen layer the reference does
    not wrap (nn.Conv1d: utils/lora.py:404-452) — base GEMM forward, dgrad backward."""

    def __init__(self, name, module, kind, *, arena=None, i_up=None, i_down=None, scale=1.0, dropout_p=0.0, pad_in=0, pad_out=0):
        self.name, self.module, self.kind = name, module, kind
        self.arena, self.i_up, self.i_down = arena, i_up, i_down
        self.scale, self.p = scale, dropout_p
        self.lora = arena is not None
        self.pad_in, self.pad_out = pad_in, pad_out      # zero-padded channel counts (0 = none): conv_in / out
        self.pk = None
        self.g_up = self.g_down = None

    def _padded(self, w, cout, cin):
        out = torch.zeros((cout, cin) + tuple(w.shape[2:]), device=w.device, dtype=torch.float32)
        out[:w.shape[0], :w.shape[1]] = w.detach().float()
        return out

    def pack(self):
        m = self.module
        w = m.weight.detach()
        bias = m.bias.detach() if m.bias is not None else None
        if w.dim() == 3:                       # Conv1d k=1 -> per-token Linear
            w = w.reshape(w.shape[0], w.shape[1])
        cout, cin = w.shape[0], w.shape[1]
        if not self.lora:
            self.w = w.reshape(cout, -1).to(BF16).contiguous()
            self.w_t = w.reshape(cout, -1).t().to(BF16).contiguous()
            self.bias = _f32(bias) if bias is not None else None
            return
        up, down = self.arena.param(self.i_up), self.arena.param(self.i_down)
        self.g_up, self.g_down = self.arena.grad(self.i_up), self.arena.grad(self.i_down)
        if self.pad_in or self.pad_out:
            pc_out, pc_in = self.pad_out or cout, self.pad_in or cin
            self._real = (up, down)
            self._up_p = torch.zeros((pc_out, 64) + tuple(up.shape[2:]), device=w.device, dtype=torch.float32)
            self._down_p = torch.zeros((64, pc_in) + tuple(down.shape[2:]), device=w.device, dtype=torch.float32)
            self._gup_p, self._gdown_p = torch.zeros_like(self._up_p), torch.zeros_like(self._down_p)
            w = self._padded(w, pc_out, pc_in)
            if bias is not None:
                b2 = torch.zeros(pc_out, device=w.device, dtype=torch.float32)
                b2[:cout] = bias.float()
                bias = b2
            self._sync_padded()
            up, down = self._up_p, self._down_p
            self.g_up, self.g_down = self._gup_p, self._gdown_p
        self.pk = _PackedLora(self.kind, w, bias, up, down, self.scale)

    def _sync_padded(self):
        up, down = self._real
        self._up_p[:up.shape[0], :up.shape[1]] = up
        self._down_p[:down.shape[0], :down.shape[1]] = down

    def refresh(self):
        """bf16 operands from the fp32 arena after an optimizer step."""
        if self.lora and self.pk is not None:
            if self.pad_in or self.pad_out:
                self._sync_padded()
            self.pk.refresh()

    def flush_padded_grads(self):
        if self.lora and (self.pad_in or self.pad_out):
            up, down = self._real
            self.arena.grad(self.i_up).add_(self._gup_p[:up.shape[0], :up.shape[1]])
            self.arena.grad(self.i_down).add_(self._gdown_p[:down.shape[0], :down.shape[1]])
            self._gup_p.zero_()
            self._gdown_p.zero_()

    # x: channels-last bf16 in the layer's point grid ([M, K] | [n, h, w, C] | [b, t, hw, C])
    def forward(self, x, training, bias_rows=None, bias_div=None, addend=None):
        """addend: the residual added to this layer's output next ([rows, Cout] bf16) — the returned y includes it."""
        if not self.lora:
            return ops.linear(x, self.w, self.bias, residual=addend), (x,)
        pk = self.pk
        y, t, mask, ms = lora_forward(pk, x, None, 1.0, bias_rows=bias_rows, bias_div=bias_div, drop_p=self.p if training else 0.0,
                                      addend=addend)
        return y, (x, t, mask, ms)

    def backward(self, saved, dy, need_dx=True):
        if not self.lora:
            return ops.linear(dy.contiguous().view(-1, dy.shape[-1]), self.w_t, None).view(*saved[0].shape) if need_dx else None
        x, t, mask, ms = saved
        dx = lora_backward(self.pk, x, t, mask, ms, dy, self.g_up, self.g_down, need_dx=need_dx)
        self.flush_padded_grads()
        return dx


class _Norm:
    def __init__(self, m):
        self.module = m
        self.w, self.b, self.eps = _f32(m.weight), _f32(m.bias), m.eps


# =============================================================================== the student
class StudentUNet:
    def __init__(self, unet: U.UNetModel, r: int = 64, dropout_p: float = 0.1, scale: float = 1.0, tconv_dropout: float = 0.1):
        dev = unet.time_embed[0].weight.device
        _require_cuda("StudentUNet", dev)      # call unet.cuda() first
        self.unet, self.device, self.r = unet, dev, r
        self.training = True
        self.tconv_p = tconv_dropout
        targets = lora_target_layers(unet)
        shapes = []
        for _, m in targets:
            rr = lora_rank(m, r)
            if isinstance(m, nn.Linear):
                shapes += [(m.out_features, rr), (rr, m.in_features)]
            else:
                k = tuple(m.kernel_size)
                shapes += [(m.out_channels, rr) + (1,) * len(k), (rr, m.in_channels) + k]
        self.arena = LoraArena(shapes, dev)
        gen = torch.Generator(device="cpu").manual_seed(0)
        self.layers = {}          # module -> _Layer
        self.layer_list = []
        for i, (name, m) in enumerate(targets):
            rr = lora_rank(m, r)
            with torch.no_grad():   # utils/lora.py:40-43: down ~ N(0, 1/r), up = 0
                self.arena.param(2 * i + 1).copy_(torch.randn(shapes[2 * i + 1], generator=gen) / rr)
            if isinstance(m, nn.Linear) or (isinstance(m, nn.Conv2d) and m.kernel_size == (1, 1)):
                kind = "linear"
            elif isinstance(m, nn.Conv2d):
                kind = "conv2d"
            else:
                kind = "conv3d"
            cin = m.in_features if isinstance(m, nn.Linear) else m.in_channels
            cout = m.out_features if isinstance(m, nn.Linear) else m.out_channels
            lay = _Layer(name, m, kind, arena=self.arena, i_up=2 * i, i_down=2 * i + 1, scale=scale, dropout_p=dropout_p,
                         pad_in=64 if cin < 64 else 0, pad_out=64 if cout < 64 else 0)
            self.layers[m] = lay
            self.layer_list.append(lay)
        self._plain = {}
        self._packed = False
        self._build()
        # first arena offset of every top-level block: once the backward has passed a block (they are visited from `out`
        # down to the middle block in reverse arena order) every gradient at or above that offset is final -> ArenaReducer.ready
        self._first_offset = {}
        for lay in self.layer_list:
            parts = lay.name.split(".")
            key = ".".join(parts[:2]) if parts[0] in ("input_blocks", "output_blocks") else parts[0]
            self._first_offset.setdefault(key, self.arena.offsets[lay.i_up])
        self.on_grads_final = None      # callable(offset): the data-parallel exchange hooks in here (dist.ArenaReducer.ready)

    # ------------------------------------------------------------------ structure
    def _L(self, m):
        if m in self.layers:
            return self.layers[m]
        if m not in self._plain:      # nn.Conv1d (init_attn proj_in / proj_out): not a LoRA target
            self._plain[m] = _Layer("", m, "linear")
        return self._plain[m]

    def _build(self):
        u = self.unet

        def block(seq):
            items = []
            for layer in seq:
                if isinstance(layer, U.ResBlock):
                    items.append(("res", self._res_struct(layer)))
                elif isinstance(layer, U.SpatialTransformer):
                    items.append(("st", self._tr_struct(layer, False)))
                elif isinstance(layer, U.TemporalTransformer):
                    items.append(("tt", self._tr_struct(layer, True)))
                elif isinstance(layer, U.Downsample):
                    items.append(("down", self._L(layer.op)))
                elif isinstance(layer, U.Upsample):
                    items.append(("up", self._L(layer.conv)))
                elif isinstance(layer, nn.Conv2d):
                    items.append(("conv_in", self._L(layer)))
                else:
                    raise TypeError(type(layer))
            return items
        self.s_input = [block(s) for s in u.input_blocks]
        self.s_init = block(u.init_attn) if u.addition_attention else None
        self.s_middle = block(u.middle_block)
        self.s_output = [block(s) for s in u.output_blocks]
        self.s_out = (_Norm(u.out[0]), self._L(u.out[2]))
        self.s_time = (self._L(u.time_embed[0]), self._L(u.time_embed[2]))
        self.s_fps = (self._L(u.fps_embedding[0]), self._L(u.fps_embedding[2])) if u.fps_cond else None
        self.s_cond = self._L(u.time_cond_proj) if u.time_cond_proj is not None else None
        self.s_motion = self._motion_struct(u)

    def _motion_struct(self, u):
        if u.motion_cond_proj is not None:
            raise NotImplementedError("StudentUNet: motion_cond_proj belongs to the v2 full fine-tune step (full_train.FullUNet)")
        return None

    # norm adjoints: the LoRA step needs dx only; full_train.FullUNet overrides these to add the affine gradients
    def _gn_bwd(self, norm, x, dy, *, rows_per_sample, silu, dx_add=None):
        return ops.groupnorm_bwd(x, dy, norm.w, norm.b, rows_per_sample=rows_per_sample, eps=norm.eps, silu=silu, dx_add=dx_add)

    def _ln_bwd(self, norm, x, dy, *, dx_add=None):
        return ops.layernorm_bwd(x, dy, norm.w, norm.eps, dx_add=dx_add)

    # temporal attention-probability export and its adjoint (motion_prior.ScoreUNet overrides these; no-ops for the training steps)
    def _probs_out(self, A, geom):
        return None

    def _probs_grad(self, A, q, k, dq, dk, geom):
        return dq, dk

    _want_input_grad = False      # True: backward() also returns d loss / d x (ScoreUNet); the training steps never need it

    def _res_struct(self, rb):
        d = dict(gn1=_Norm(rb.in_layers[0]), conv1=self._L(rb.in_layers[2]), emb=self._L(rb.emb_layers[1]),
                 gn2=_Norm(rb.out_layers[0]), conv2=self._L(rb.out_layers[3]), cout=rb.out_channels,
                 skip=None if isinstance(rb.skip_connection, nn.Identity) else self._L(rb.skip_connection), tconv=None)
        if rb.use_temporal_conv:
            d["tconv"] = [(_Norm(getattr(rb.temopral_conv, f"conv{i}")[0]), self._L(getattr(rb.temopral_conv, f"conv{i}")[-1]))
                          for i in (1, 2, 3, 4)]
        return d

    def _tr_struct(self, m, temporal):
        blk = m.transformer_blocks[0]

        def attn(a):
            return dict(q=self._L(a.to_q), k=self._L(a.to_k), v=self._L(a.to_v), o=self._L(a.to_out[0]), heads=a.heads, scale=a.scale)
        return dict(temporal=temporal, gn=_Norm(m.norm), proj_in=self._L(m.proj_in), proj_out=self._L(m.proj_out),
                    a1=attn(blk.attn1), a2=attn(blk.attn2), ln=[_Norm(blk.norm1), _Norm(blk.norm2), _Norm(blk.norm3)],
                    ff1=self._L(blk.ff.net[0].proj), ff2=self._L(blk.ff.net[2]))

    def pack(self):
        for lay in list(self.layers.values()) + list(self._plain.values()):
            lay.pack()
        half = self.unet.model_channels // 2
        self.freqs = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32) / half).to(self.device)
        self._packed = True

    def refresh(self):
        """Re-derive the bf16 LoRA operands after `arena.adamw_step` / `arena.load_list` (in place; ~2 300 small device ops).
        `graph_refresh()` captures them once as a CUDA graph, after which this is a single replay."""
        if getattr(self, "_refresh_graph", None) is not None:
            self._refresh_graph.replay()
            return
        for lay in self.layer_list:
            lay.refresh()

    def graph_refresh(self):
        if not self._packed:
            self.pack()
        self._refresh_graph = None
        self.refresh()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            for lay in self.layer_list:
                lay.refresh()
        self._refresh_graph = g

    def train(self, mode=True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)

    # ------------------------------------------------------------------ embeddings
    def _emb_fwd(self, timesteps, fps, timestep_cond, bsz, motion_cond=None):
        if motion_cond is not None:
            raise NotImplementedError("StudentUNet: motion_cond belongs to the v2 full fine-tune step (full_train.FullUNet)")
        tr, dev = self.training, self.device
        ctx = {}
        t = timesteps.to(device=dev, dtype=torch.float32).reshape(-1)
        if t.numel() == 1 and bsz > 1:
            t = t.expand(bsz).contiguous()
        t_emb = ops.sinusoidal_embedding(t, self.freqs, round_bf16=True)
        if timestep_cond is not None:
            cond, ctx["cond"] = self.s_cond.forward(timestep_cond.to(device=dev, dtype=BF16).contiguous(), tr)
            t_emb = t_emb + cond.float()      # [B, 320] fp32 add: plumbing
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
        self._emb_act = ops.silu(emb)                 # SiLU(emb): the shared input of the 22+ emb_layers projections
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
            self.s_cond.backward(ctx["cond"], d_t, need_dx=False)

    def _acc_emb(self, g):
        self._d_emb_act = g if self._d_emb_act is None else ops.add(self._d_emb_act, g)

    # ------------------------------------------------------------------ ResBlock
    def _res_fwd(self, R, x, geom):
        b, t, hh, ww = geom
        nf, hw = b * t, hh * ww
        tr = self.training
        cin, cout = x.shape[-1], R["cout"]
        x2 = x.view(-1, cin)
        c = dict(x=x2, geom=geom)
        hn = ops.groupnorm(x2, R["gn1"].w, R["gn1"].b, rows_per_sample=hw, eps=R["gn1"].eps, silu=True)
        rb, c["emb"] = R["emb"].forward(self._emb_act, tr)                       # [B, cout]
        bias_rows = (rb.float() + R["conv1"].pk.bias).contiguous()               # conv bias + per-sample embedding row
        h, c["conv1"] = R["conv1"].forward(hn.view(nf, hh, ww, cin), tr, bias_rows=bias_rows, bias_div=t)
        c["h"] = h.view(-1, cout)
        hn2 = ops.groupnorm(c["h"], R["gn2"].w, R["gn2"].b, rows_per_sample=hw, eps=R["gn2"].eps, silu=True)
        if R["skip"] is None:
            skip = x2
        else:
            skip, c["skip"] = R["skip"].forward(x2, tr)
        h2, c["conv2"] = R["conv2"].forward(hn2.view(nf, hh, ww, cout), tr, addend=skip)   # h + out_layers(...) (openaimodel3d.py:246-250)
        h2 = h2.view(-1, cout)
        if R["tconv"] is not None:
            y = h2
            c["tc"] = []
            for i, (gn, conv) in enumerate(R["tconv"]):
                yn = ops.groupnorm(y, gn.w, gn.b, rows_per_sample=t * hw, eps=gn.eps, silu=True)
                mask, ms = None, 1.0
                if tr and i > 0 and self.tconv_p > 0.0:                                            # openaimodel3d.py:280-296
                    yn, mask = ops.dropout_scale(yn, self.tconv_p)
                    ms = 1.0 / (1.0 - self.tconv_p)
                y_in = y
                y, sv = conv.forward(yn.view(b, t, hw, cout), tr)
                y = y.view(-1, cout)
                c["tc"].append((y_in, sv, mask, ms))
            h2 = ops.add(h2, y)
        return h2.view(nf, hh, ww, cout), c

    def _res_bwd(self, R, c, dout):
        b, t, hh, ww = c["geom"]
        nf, hw = b * t, hh * ww
        cout = R["cout"]
        d_h2 = dout.reshape(-1, cout)
        if R["tconv"] is not None:
            dy = d_h2
            for i in (3, 2, 1, 0):
                gn, conv = R["tconv"][i]
                y_in, sv, mask, ms = c["tc"][i]
                d_yn = conv.backward(sv, dy.view(b, t, hw, cout)).view(-1, cout)
                if mask is not None:
                    d_yn = ops.scale_mask(d_yn, ms, mask)
                dy = self._gn_bwd(gn, y_in, d_yn, rows_per_sample=t * hw, silu=True, dx_add=d_h2 if i == 0 else None)
            d_h2 = dy
        d_hn2 = R["conv2"].backward(c["conv2"], d_h2.view(nf, hh, ww, cout)).view(-1, cout)
        d_skip = d_h2 if R["skip"] is None else R["skip"].backward(c["skip"], d_h2)
        d_h = self._gn_bwd(R["gn2"], c["h"], d_hn2, rows_per_sample=hw, silu=True)
        d_rb = ops.colsum_samples(d_h, t * hw)                                   # fp32 [B, cout]
        self._acc_emb(R["emb"].backward(c["emb"], d_rb.to(BF16)))
        cin = c["x"].shape[-1]
        d_hn = R["conv1"].backward(c["conv1"], d_h.view(nf, hh, ww, cout)).view(-1, cin)
        dx = self._gn_bwd(R["gn1"], c["x"], d_hn, rows_per_sample=hw, silu=True, dx_add=d_skip)
        return dx.view(nf, hh, ww, cin)

    # ------------------------------------------------------------------ transformers
    def _attn_fwd(self, A, xn, kv_src, geom, temporal, c, skip=None):
        """xn: normalised tokens [rows, C]; kv_src: tokens the keys / values are projected from ([rows, C] or the
        frame-repeated text context); skip: the residual `x + attn(norm(x))` (attention.py:276-281), folded into to_out."""
        b, t, hh, ww = geom
        hw = hh * ww
        tr = self.training
        heads, scale = A["heads"], A["scale"]
        q, c["q"] = A["q"].forward(xn, tr)
        k, c["k"] = A["k"].forward(kv_src, tr)
        v, c["v"] = A["v"].forward(kv_src, tr)
        inner = q.shape[-1]
        if temporal:
            att = ops.attention_temporal(q, k, v, b=b, t=t, hw=hw, heads=heads, scale=scale, probs=self._probs_out(A, geom))
            c["att"] = (q, k, v)
        else:
            bt = b * t
            lk = k.shape[0] // bt
            lse2 = torch.empty((bt, heads, hw), device=q.device, dtype=torch.float32)
            q3, k3, v3 = q.view(bt, hw, inner), k.view(bt, lk, inner), v.view(bt, lk, inner)
            o3 = ops.attention(q3, k3, v3, heads=heads, scale=scale, lse2=lse2)
            att = o3.view(-1, inner)
            c["att"] = (q3, k3, v3, o3, lse2)
        out, c["o"] = A["o"].forward(att, tr, addend=skip)
        return out

    def _attn_bwd(self, A, c, d_out, geom, temporal, self_attn):
        """-> (d_xn, d_kv_src or None)"""
        b, t, hh, ww = geom
        hw = hh * ww
        d_att = A["o"].backward(c["o"], d_out)
        if temporal:
            q, k, v = c["att"]
            dq, dk, dv = ops.attention_temporal_bwd(q, k, v, d_att, b=b, t=t, hw=hw, heads=A["heads"], scale=A["scale"])
            dq, dk = self._probs_grad(A, q, k, dq, dk, geom)
        else:
            q3, k3, v3, o3, lse2 = c["att"]
            dq, dk, dv = ops.attention_bwd(q3, k3, v3, o3, d_att.view(*o3.shape), lse2, heads=A["heads"], scale=A["scale"])
            dq, dk, dv = dq.view(-1, dq.shape[-1]), dk.view(-1, dk.shape[-1]), dv.view(-1, dv.shape[-1])
        d_xn = A["q"].backward(c["q"], dq)
        d_k = A["k"].backward(c["k"], dk, need_dx=self_attn)
        d_v = A["v"].backward(c["v"], dv, need_dx=self_attn)
        if self_attn:
            d_xn = ops.add(ops.add(d_xn, d_k), d_v)
        return d_xn

    def _tr_fwd(self, T, h, geom, ctx_rows):
        b, t, hh, ww = geom
        hw = hh * ww
        tr = self.training
        temporal = T["temporal"]
        ch = h.shape[-1]
        x_in = h.view(-1, ch)
        c = dict(x_in=x_in, geom=geom)
        rps = hw * (t if temporal else 1)
        xn = ops.groupnorm(x_in, T["gn"].w, T["gn"].b, rows_per_sample=rps, eps=T["gn"].eps, silu=False)
        x0, c["pin"] = T["proj_in"].forward(xn, tr)
        ln = T["ln"]
        c["x0"] = x0
        n1 = ops.layernorm(x0, ln[0].w, ln[0].b, ln[0].eps)
        c["a1"] = {}
        x1 = self._attn_fwd(T["a1"], n1, n1, geom, temporal, c["a1"], skip=x0)
        c["x1"] = x1
        n2 = ops.layernorm(x1, ln[1].w, ln[1].b, ln[1].eps)
        c["a2"] = {}
        kv_src = n2 if temporal else ctx_rows
        x2 = self._attn_fwd(T["a2"], n2, kv_src, geom, temporal, c["a2"], skip=x1)
        c["x2"] = x2
        n3 = ops.layernorm(x2, ln[2].w, ln[2].b, ln[2].eps)
        pre, c["ff1"] = T["ff1"].forward(n3, tr)
        c["pre"] = pre
        g = ops.geglu(pre)
        x3, c["ff2"] = T["ff2"].forward(g, tr, addend=x2)
        out, c["pout"] = T["proj_out"].forward(x3, tr, addend=x_in)
        return out.view(*h.shape), c

    def _tr_bwd(self, T, c, dout):
        geom = c["geom"]
        b, t, hh, ww = geom
        hw = hh * ww
        temporal = T["temporal"]
        ln = T["ln"]
        d_out = dout.reshape(-1, dout.shape[-1])
        d_x3 = T["proj_out"].backward(c["pout"], d_out)
        d_g = T["ff2"].backward(c["ff2"], d_x3)
        d_pre = ops.geglu(c["pre"], d_g)
        d_n3 = T["ff1"].backward(c["ff1"], d_pre)
        d_x2 = self._ln_bwd(ln[2], c["x2"], d_n3, dx_add=d_x3)
        d_n2 = self._attn_bwd(T["a2"], c["a2"], d_x2, geom, temporal, self_attn=temporal)
        d_x1 = self._ln_bwd(ln[1], c["x1"], d_n2, dx_add=d_x2)
        d_n1 = self._attn_bwd(T["a1"], c["a1"], d_x1, geom, temporal, self_attn=True)
        d_x0 = self._ln_bwd(ln[0], c["x0"], d_n1, dx_add=d_x1)
        d_xn = T["proj_in"].backward(c["pin"], d_x0)
        rps = hw * (t if temporal else 1)
        dx = self._gn_bwd(T["gn"], c["x_in"], d_xn, rows_per_sample=rps, silu=False, dx_add=d_out)
        return dx.view(*dout.shape)

    # ------------------------------------------------------------------ sequences
    def _seq_fwd(self, items, h, geom, ctx_rows, tape):
        b, t, hh, ww = geom
        tr = self.training
        for kind, S in items:
            if kind == "res":
                h, c = self._res_fwd(S, h, geom)
            elif kind in ("st", "tt"):
                h, c = self._tr_fwd(S, h, geom, ctx_rows)
            elif kind == "down":       # stride-2 conv = the stride-1 conv (with its LoRA branch) sampled at even positions
                y, sv = S.forward(h, tr)
                h = ops.resample2x(y, "sub")
                c = sv
                hh, ww = hh // 2, ww // 2
                geom = (b, t, hh, ww)
            elif kind == "up":
                y, c = S.forward(ops.upsample_nearest2x(h), tr)
                h = y
                hh, ww = hh * 2, ww * 2
                geom = (b, t, hh, ww)
            elif kind == "conv_in":
                h, c = S.forward(h, tr)
            tape.append((kind, S, c))
        return h, geom

    def _seq_bwd(self, tape, dh):
        for kind, S, c in reversed(tape):
            if kind == "res":
                dh = self._res_bwd(S, c, dh)
            elif kind in ("st", "tt"):
                dh = self._tr_bwd(S, c, dh)
            elif kind == "down":
                dh = S.backward(c, ops.resample2x(dh.contiguous(), "stuff"))
            elif kind == "up":
                dh = ops.resample2x(S.backward(c, dh), "pool")
            elif kind == "conv_in":
                dh = S.backward(c, dh, need_dx=self._want_input_grad)
        return dh

    # ------------------------------------------------------------------ forward / backward
    def forward(self, x, timesteps, context=None, fps=16, timestep_cond=None, motion_cond=None, **kwargs):
        _require_cuda(type(self).__name__ + " input", x.device)
        if not self._packed:
            self.pack()
        if self.training:
            ops.dropout_advance(x.device)   # fresh in-kernel dropout masks for this forward (capturable: a device-side add)
        u = self.unet
        b, cin, t, hh, ww = x.shape
        self._emb_fwd(timesteps, fps, timestep_cond, b, motion_cond)
        ctx_rows = None
        if context is not None:    # the reference repeats the text context per frame before to_k / to_v (openaimodel3d.py:710)
            ctx_rows = context.to(device=self.device, dtype=BF16).repeat_interleave(t, 0).reshape(-1, context.shape[-1]).contiguous()
        h = ops.bcthw_to_frames_pad(x, 64)
        geom = (b, t, hh, ww)
        self._tapes_in, self._tape_mid, self._tapes_out = [], [], []
        hs = []
        for i, items in enumerate(self.s_input):
            tape = []
            h, geom = self._seq_fwd(items, h, geom, ctx_rows, tape)
            if i == 0 and self.s_init is not None:
                h, geom = self._seq_fwd(self.s_init, h, geom, ctx_rows, tape)
            self._tapes_in.append(tape)
            hs.append(h)
        h, geom = self._seq_fwd(self.s_middle, h, geom, ctx_rows, self._tape_mid)
        self._skip_ch = []
        for items in self.s_output:
            skip = hs.pop()
            self._skip_ch.append((h.shape[-1], skip.shape[-1]))
            h = ops.concat_channels(h, skip)
            tape = []
            h, geom = self._seq_fwd(items, h, geom, ctx_rows, tape)
            self._tapes_out.append(tape)
        gn, conv = self.s_out
        c = h.shape[-1]
        self._out_ctx = dict(h=h.view(-1, c), geom=geom)
        hn = ops.groupnorm(h.view(-1, c), gn.w, gn.b, rows_per_sample=geom[2] * geom[3], eps=gn.eps, silu=True)
        y, self._out_ctx["conv"] = conv.forward(hn.view(b * t, geom[2], geom[3], c), self.training)
        return ops.frames_to_bcthw(y, b, u.out_channels, x.dtype)

    __call__ = forward

    def detach_tapes(self):
        """Take the saved activations of the last forward out of the object (so that another, gradient-free forward can run
        before the backward: the distillation step's target prediction); hand them back with restore_tapes."""
        saved = (self._tapes_in, self._tape_mid, self._tapes_out, self._out_ctx, self._emb_ctx, self._emb_act, self._skip_ch)
        self._tapes_in = self._tape_mid = self._tapes_out = self._out_ctx = self._emb_ctx = self._emb_act = None
        return saved

    def restore_tapes(self, saved):
        self._tapes_in, self._tape_mid, self._tapes_out, self._out_ctx, self._emb_ctx, self._emb_act, self._skip_ch = saved
        self._d_emb_act = None

    def backward(self, d_out):
        """d_out: gradient of the loss w.r.t. the forward's output [B, 4, T, H, W].  Accumulates into arena.grads."""
        gn, conv = self.s_out
        oc = self._out_ctx
        b, t, hh, ww = oc["geom"]
        dy = ops.bcthw_to_frames_pad(d_out, 64)
        c = oc["h"].shape[-1]
        d_hn = conv.backward(oc["conv"], dy).view(-1, c)
        dh = self._gn_bwd(gn, oc["h"], d_hn, rows_per_sample=hh * ww, silu=True)
        d_skips = []
        n_out = len(self._tapes_out)
        for j, (tape, (c_h, c_s)) in enumerate(zip(reversed(self._tapes_out), reversed(self._skip_ch))):
            dcat = self._seq_bwd(tape, dh)
            if self.on_grads_final is not None:
                self.on_grads_final(self._first_offset[f"output_blocks.{n_out - 1 - j}"])
            d2 = dcat.view(-1, c_h + c_s)
            dh = d2[:, :c_h].contiguous().view(*dcat.shape[:-1], c_h)
            d_skips.append(d2[:, c_h:])
        dh = self._seq_bwd(self._tape_mid, dh)
        if self.on_grads_final is not None:     # (init_attn sits between input_blocks and middle_block in the arena but runs with
            self.on_grads_final(self._first_offset["middle_block"])   # input block 0: nothing below the middle block is final yet)
        for tape in reversed(self._tapes_in):
            ds = d_skips.pop()
            dh = ops.add(dh.reshape(-1, dh.shape[-1]), ds).view(*dh.shape)
            dh = self._seq_bwd(tape, dh)
        self._emb_bwd()
        if self.on_grads_final is not None:
            self.on_grads_final(0)
        # release the saved activations
        self._tapes_in = self._tape_mid = self._tapes_out = self._out_ctx = self._emb_ctx = None
        return dh          # None unless _want_input_grad: then d loss / d x as frames [B*T, H, W, 64] (the first in_channels are real)
