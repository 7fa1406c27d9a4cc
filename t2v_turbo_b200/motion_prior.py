"""The motion-prior score of the v2 preprocessing on B200 (motion_prior_sample.py:40-84, utils/common_utils.py:446-478; consumed by
preprocess_scripts/preprocess_with_motion_prior.py:357-401, which stores it as the `score` of every latent-dataset sample):

    probs_ref = temporal attention probabilities of the teacher on the EXAMPLE latent             [no grad]
    eps, probs = teacher(latents, ts, context)            with the probabilities of the same layers
    loss  = temp_loss_scale * mean_layers( 100 * mse(probs_ref[top-1 mask], probs[top-1 mask]) )   compute_temp_loss
    score = d loss / d latents                                                                    torch.autograd.grad in the reference

The loss never touches the UNet's OUTPUT: its gradient enters through the exported attention probabilities (`record_attn_probs`,
attention.py:124-126 — the decoder's temporal `attn1` modules, output_blocks.3-11 in the VC2 UNet) and leaves through the latents.
`ScoreUNet` is the frozen-teacher view for that: `train_unet.StudentUNet`'s traversal and input-gradient kernels with

    * plain frozen GEMM layers (forward implicit GEMM, backward dgrad only: no weight gradient of any kind is formed),
    * the probabilities of the recording layers exported by t2v_attn_short_fwd in the forward,
    * their gradient turned into (dq, dk) by t2v_attn_short_probs_bwd in the backward and added to the value path's,
    * the input gradient carried through conv_in (`_want_input_grad`).

The loss itself is a few torch reductions on the exported tensors (host-side plumbing, fp32).

Status: composition verified on CPU against the UNMODIFIED reference's autograd (tests/golden/motion_score_small.pt)
through the kernel-contract restatements; the new kernel runs under the host emulation; never run on a GPU.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import ops
from . import unet as U
from .lora_train import _base_op
from .train_unet import StudentUNet, _require_cuda

BF16 = torch.bfloat16


class _FrozenLayer:
    """A frozen GEMM layer of the teacher: bf16 forward operand, dgrad operand (transposed, taps reversed), fp32 bias.  The 4-channel
    conv_in / out run zero-padded to 64 channels as in the training views."""
    lora = False

    def __init__(self, name, module, kind, pad_in=0, pad_out=0):
        self.name, self.module, self.kind = name, module, kind
        w = module.weight.detach()
        self.cout, self.cin = w.shape[0], w.shape[1]
        self.pc_out, self.pc_in = pad_out or self.cout, pad_in or self.cin
        self.pk = self
        self.w = self.w_t = self.bias = None

    def pack(self):
        m = self.module
        w = m.weight.detach().float()
        if self.kind == "linear":
            w = w.reshape(self.cout, -1)
        if self.pc_out != self.cout or self.pc_in != self.cin:
            wp = torch.zeros((self.pc_out, self.pc_in) + tuple(w.shape[2:]), device=w.device, dtype=torch.float32)
            wp[:self.cout, :self.cin] = w
            w = wp
        if self.kind == "linear":
            self.w, self.w_t = w.to(BF16).contiguous(), w.t().to(BF16).contiguous()
        else:
            sp = tuple(range(2, w.dim()))
            self.w = ops.pack_conv_weight(w)
            self.w_t = ops.pack_conv_weight(w.transpose(0, 1).flip(sp).contiguous())
        if m.bias is not None:
            b = torch.zeros(self.pc_out, device=w.device, dtype=torch.float32)
            b[:self.cout] = m.bias.detach().float()
            self.bias = b

    def refresh(self):
        pass

    def forward(self, x, training, bias_rows=None, bias_div=None, addend=None):
        bias = self.bias if bias_rows is None else bias_rows
        res = addend.view(*x.shape[:-1], self.pc_out) if addend is not None else None
        return _base_op(self.kind, x, self.w, bias, residual=res, bias_div=bias_div), (x,)

    def backward(self, saved, dy, need_dx=True):
        if not need_dx:
            return None
        (x,) = saved
        return _base_op(self.kind, dy.contiguous().view(*x.shape[:-1], self.pc_out), self.w_t, None)


class ScoreUNet(StudentUNet):
    """Frozen-teacher view of a `UNetModel` built with `record_attn_probs=True`: forward -> (eps, {layer name: probabilities}),
    backward(d_probs) -> d loss / d latents."""
    _want_input_grad = True

    def __init__(self, unet: U.UNetModel):
        dev = unet.time_embed[0].weight.device
        _require_cuda("ScoreUNet", dev)
        self.unet, self.device = unet, dev
        self.training = False
        self.tconv_p = 0.0
        self._mod_name = {m: n for n, m in unet.named_modules()}
        self.layers, self.layer_list, self._plain = {}, [], {}
        self._packed = False
        self.on_grads_final = None
        self._first_offset = {}
        self.probs, self.d_probs = {}, {}
        self._build()
        self.probe_names = [n for n, m in unet.named_modules() if isinstance(m, U.CrossAttention) and getattr(m, "record_attn_probs", False)]
        if not self.probe_names:
            raise ValueError("ScoreUNet: the UNet records no attention probabilities (build it with record_attn_probs=True)")

    # ------------------------------------------------------------------ structure
    def _L(self, m):
        lay = self.layers.get(m)
        if lay is None:
            if isinstance(m, (nn.Linear, nn.Conv1d)) or (isinstance(m, nn.Conv2d) and m.kernel_size == (1, 1)):
                kind = "linear"
            elif isinstance(m, nn.Conv2d):
                kind = "conv2d"
            else:
                kind = "conv3d"
            cin = m.in_features if isinstance(m, nn.Linear) else m.in_channels
            cout = m.out_features if isinstance(m, nn.Linear) else m.out_channels
            lay = _FrozenLayer(self._mod_name[m], m, kind, pad_in=64 if cin < 64 else 0, pad_out=64 if cout < 64 else 0)
            self.layers[m] = lay
            self.layer_list.append(lay)
        return lay

    def _motion_struct(self, u):          # the teacher of the preprocessing is the plain VC2 UNet; a motion-conditioned one has
        return None                       # its extra projections simply unused here (motion_cond is never passed)

    def _tr_struct(self, m, temporal):
        d = super()._tr_struct(m, temporal)
        a1 = m.transformer_blocks[0].attn1
        if temporal and getattr(a1, "record_attn_probs", False):
            d["a1"]["probe"] = self._mod_name[a1]
        return d

    def pack(self):
        for lay in self.layer_list:
            lay.pack()
        half = self.unet.model_channels // 2
        self.freqs = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32) / half).to(self.device)
        self._packed = True

    # ------------------------------------------------------------------ hooks of the traversal
    def _probs_out(self, A, geom):
        name = A.get("probe")
        if name is None:
            return None
        b, t, hh, ww = geom
        buf = torch.empty((b * hh * ww * A["heads"], t, t), device=self.device, dtype=torch.float32)
        self.probs[name] = buf
        return buf

    def _probs_grad(self, A, q, k, dq, dk, geom):
        dp = self.d_probs.get(A.get("probe"))
        if dp is None:
            return dq, dk
        b, t, hh, ww = geom
        dq2, dk2 = ops.attention_temporal_probs_bwd(q, k, dp.contiguous(), b=b, t=t, hw=hh * ww, heads=A["heads"], scale=A["scale"])
        return ops.add(dq, dq2), ops.add(dk, dk2)

    def _emb_bwd(self):                   # nothing trains: the embedding path's gradient has no consumer
        pass

    def _acc_emb(self, g):
        pass

    # ------------------------------------------------------------------ API
    def forward(self, x, timesteps, context=None, fps=16, timestep_cond=None, **kwargs):
        self.probs = {}
        y = super().forward(x, timesteps, context=context, fps=fps, timestep_cond=timestep_cond)
        self._x_meta = (x.shape[0], x.shape[1])
        return y, dict(self.probs)

    __call__ = forward

    def backward(self, d_probs, d_out=None):
        """d_probs: {layer name: fp32 gradient w.r.t. that layer's exported probabilities}; d_out: optional gradient w.r.t. the
        output (zero for the motion-prior loss).  -> d loss / d x, fp32 [B, C, T, H, W]."""
        b, c = self._x_meta
        self.d_probs = d_probs
        oc = self._out_ctx
        bb, t, hh, ww = oc["geom"]
        if d_out is None:
            d_out = torch.zeros((bb, self.unet.out_channels, t, hh, ww), device=self.device, dtype=torch.float32)
        dh = super().backward(d_out)
        self.d_probs = {}
        return ops.frames_to_bcthw(dh.contiguous(), b, c, torch.float32)


def temp_loss_and_grad(probs, probs_example, temp_loss_scale=1.0, rank_k=1):
    """temp_loss_scale * compute_temp_loss(probs, probs_example) (utils/common_utils.py:446-478) and its gradient w.r.t. every tensor
    of `probs`: per layer, the mean squared error on the positions of the top-`rank_k` entries of the EXAMPLE's probabilities along the
    key axis, times 100, averaged over layers.  A few fp32 torch reductions on the exported tensors (plumbing, not the hot path)."""
    leaves = {n: p.detach().float().requires_grad_(True) for n, p in probs.items()}
    losses = []
    for n, gen in leaves.items():
        ref = probs_example[n].detach().float()
        if rank_k > ref.shape[-1]:
            raise ValueError("The value of rank_k cannot be larger than the number of frames")
        _, idx = torch.sort(ref, dim=-1)
        mask = torch.zeros_like(ref, dtype=torch.bool)
        mask.scatter_(-1, idx[..., -rank_k:], True)
        losses.append(torch.nn.functional.mse_loss(ref[mask], gen[mask]))
    loss = temp_loss_scale * (torch.stack(losses) * 100).mean()
    loss.backward()
    return loss.detach(), {n: g.grad for n, g in leaves.items()}


def get_motion_prior_score(view: ScoreUNet, latents, ts, example_latent, original_context, inference_context, temp_loss_scale):
    """motion_prior_sample.py:59-84 -> (score = d loss / d latents, cond_teacher_output).  The contexts are the reference's dicts
    ({"context": text embedding, "fps": 16})."""
    _, probs_example = view(example_latent, ts, **original_context)
    view.detach_tapes()
    cond_teacher_output, probs = view(latents, ts, **inference_context)
    _, d_probs = temp_loss_and_grad(probs, probs_example, temp_loss_scale)
    score = view.backward(d_probs)
    return score.to(latents.dtype), cond_teacher_output


def reverse_ddim_loop(unet_fn, latents, context, solver, num_inference_steps):
    """motion_prior_sample.py:27-37: DDIM inversion of a clean latent, one step per solver timestep; returns every intermediate latent.
    unet_fn(latents, ts, **context) -> eps (no grad: the fused inference `UNetModel`, or a ScoreUNet's forward)."""
    out = []
    for i in range(num_inference_steps):
        ts = solver.ddim_timesteps[torch.tensor([i])].long()
        eps = unet_fn(latents, ts.to(latents.device), **context)
        latents = solver.ddim_reverse_step(latents, eps.to(latents.dtype), ts)
        out.append(latents)
    return out


def preprocess_sample(view: ScoreUNet, scheduler, solver, latents, prompt_emb, uncond_emb, *, index, noise, temp_loss_scale, fps=16,
                      unet_fn=None):
    """The device work of preprocess_scripts/preprocess_with_motion_prior.py:326-401 for ONE video (the script's batch size): from the
    scaled VAE latent [1, C, T, h, w] and the text embeddings to the record of the v2 latent dataset — z_t = add_noise(latents, noise,
    t), the DDIM-inverted example latents (index + 1 teacher forwards), the unconditional teacher output, and the motion-prior score with
    the conditional teacher output — ready for `formats.dumps_v2_sample(**record)`.  unet_fn: optional faster no-grad forward (the fused
    inference UNetModel of the same weights); default: the ScoreUNet's own forward."""
    if unet_fn is None:
        def unet_fn(x, ts, **ctx):
            y, _ = view(x, ts, **ctx)
            view.detach_tapes()
            return y
    index = torch.as_tensor(index).reshape(1).long()
    start_t = solver.ddim_timesteps[index.cpu()]
    ctx = {"context": prompt_emb, "fps": fps}
    z_t = scheduler.add_noise(latents, noise, start_t.to(latents.device))                          # :341-343
    inter = reverse_ddim_loop(unet_fn, latents, ctx, solver, int(index.item()) + 1)                # :346-353
    z_example = inter[-1]
    z_example_prev = inter[-2] if int(index.item()) > 0 else latents
    uncond_out = unet_fn(z_t, start_t.to(latents.device), context=uncond_emb, fps=fps)             # :355
    score, cond_out = get_motion_prior_score(view, z_t, start_t.to(latents.device), z_example, ctx, ctx, temp_loss_scale)   # :358-368
    return dict(index=index[0], z_t=z_t[0], cond_teacher_out=cond_out[0], uncond_teacher_out=uncond_out[0], score=score[0],
                z_example=z_example[0], z_example_prev=z_example_prev[0], prompt_emb=prompt_emb[0])
