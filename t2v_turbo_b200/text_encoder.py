"""The step BEFORE the denoising path: `FrozenOpenCLIPEmbedder` (lvdm/modules/encoders/condition.py:
209-283) — the text tower of OpenCLIP ViT-H-14 (24 pre-LN transformer blocks, width 1024, 16 heads of 64, 77 tokens, causal
attention mask), layer = "penultimate" (the last block is skipped) followed by `ln_final` — on the B200 kernels.

    x = token_embedding[tokens] + positional_embedding                         t2v_embedding_gather
    for the first 23 blocks:  x += out_proj(attn(ln_1(x)))  ;  x += c_proj(gelu(c_fc(ln_2(x))))
         ln_1 / ln_2  t2v_layernorm;  in_proj (fused q|k|v), out_proj (+residual), c_fc (+erf-GELU epilogue), c_proj (+residual):
         tcgen05 GEMMs;  attention: t2v_attn_fwd with causal = 1 (heads of 64, q/k/v read in place from the in_proj output)
    z = ln_final(x)                                                              [B, 77, 1024]

State-dict keys are open_clip's (`model.token_embedding.weight`, `model.positional_embedding`,
`model.transformer.resblocks.N.{ln_1, attn.in_proj_weight, attn.in_proj_bias, attn.out_proj, ln_2, mlp.c_fc, mlp.c_proj}`,
`model.ln_final`), i.e. the `cond_stage_model.*` entries of the VideoCrafter2 checkpoint load as they are.  The BPE
tokenizer (`open_clip.tokenize`) is host string processing and not part of this package: pass token ids, or a tokenizer
callable.  PARITY UNPINNED: open_clip is not installed here; the oracle (oracle/text_oracle.py) restates its published
text transformer (nn.MultiheadAttention blocks, additive -inf causal mask, erf-GELU MLP).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops

BF16 = torch.bfloat16


class _Block(nn.Module):
    def __init__(self, width, mlp_ratio=4):
        super().__init__()
        self.ln_1 = nn.LayerNorm(width)
        self.attn = nn.Module()
        self.attn.in_proj_weight = nn.Parameter(torch.empty(3 * width, width))
        self.attn.in_proj_bias = nn.Parameter(torch.zeros(3 * width))
        self.attn.out_proj = nn.Linear(width, width)
        self.ln_2 = nn.LayerNorm(width)
        self.mlp = nn.Module()
        self.mlp.c_fc = nn.Linear(width, width * mlp_ratio)
        self.mlp.c_proj = nn.Linear(width * mlp_ratio, width)
        nn.init.normal_(self.attn.in_proj_weight, std=width ** -0.5)


class FrozenOpenCLIPEmbedder(nn.Module):
    LAYERS = ["last", "penultimate"]

    def __init__(self, arch="ViT-H-14", version="laion2b_s32b_b79k", device="cuda", max_length=77, freeze=True, layer="last",
                 width=1024, layers=24, heads=16, vocab_size=49408, tokenizer=None):
        super().__init__()
        assert layer in self.LAYERS and width == heads * 64, "heads of 64 (ViT-H-14 text tower: 1024 = 16 x 64)"
        self.max_length, self.layer, self.heads, self.width = max_length, layer, heads, width
        self.layer_idx = 0 if layer == "last" else 1
        self.tokenizer = tokenizer
        m = self.model = nn.Module()
        m.token_embedding = nn.Embedding(vocab_size, width)
        m.positional_embedding = nn.Parameter(torch.empty(max_length, width).normal_(std=0.01))
        m.transformer = nn.Module()
        m.transformer.resblocks = nn.ModuleList([_Block(width) for _ in range(layers)])
        m.ln_final = nn.LayerNorm(width)
        m.text_projection = nn.Parameter(torch.empty(width, width).normal_(std=width ** -0.5))   # unused by the embedder
        m.logit_scale = nn.Parameter(torch.ones([]) * 2.6593)
        self._packed = None
        if freeze:
            self.freeze()

    def freeze(self):
        self.eval()
        for p in self.parameters():
            p.requires_grad = False

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._packed = None
        return super().load_state_dict(*a, **k)

    @torch.no_grad()
    def pack(self):
        m = self.model
        if m.positional_embedding.device.type != "cuda":
            raise RuntimeError("FrozenOpenCLIPEmbedder(B200) runs on a CUDA device only (no CPU fallback)")
        f32 = lambda t: t.detach().float().contiguous()
        b16 = lambda t: t.detach().to(BF16).contiguous()
        P = dict(tok=m.token_embedding.weight.detach().float().contiguous(), pos=f32(m.positional_embedding), blocks=[],
                 ln_final=(f32(m.ln_final.weight), f32(m.ln_final.bias), m.ln_final.eps))
        n = len(m.transformer.resblocks) - self.layer_idx
        for blk in list(m.transformer.resblocks)[:n]:
            P["blocks"].append(dict(
                ln1=(f32(blk.ln_1.weight), f32(blk.ln_1.bias), blk.ln_1.eps), ln2=(f32(blk.ln_2.weight), f32(blk.ln_2.bias), blk.ln_2.eps),
                w_qkv=b16(blk.attn.in_proj_weight), b_qkv=f32(blk.attn.in_proj_bias),
                w_o=b16(blk.attn.out_proj.weight), b_o=f32(blk.attn.out_proj.bias),
                w_fc=b16(blk.mlp.c_fc.weight), b_fc=f32(blk.mlp.c_fc.bias), w_pr=b16(blk.mlp.c_proj.weight), b_pr=f32(blk.mlp.c_proj.bias)))
        self._packed = P
        return self

    @torch.no_grad()
    def encode_with_transformer(self, tokens):
        """tokens: int64 [B, 77] (open_clip.tokenize output) -> [B, 77, width] in the dtype of the parameters."""
        if not tokens.is_cuda:
            raise RuntimeError("FrozenOpenCLIPEmbedder(B200): tokens must be a CUDA tensor (no CPU fallback)")
        if self._packed is None:
            self.pack()
        P = self._packed
        b, n = tokens.shape
        w, h = self.width, self.heads
        x = ops.embedding_gather(P["tok"], tokens.to(torch.int64), P["pos"][:n].contiguous())          # [B*n, W] bf16
        for blk in P["blocks"]:
            y = ops.layernorm(x, blk["ln1"][0], blk["ln1"][1], blk["ln1"][2])
            qkv = ops.linear(y, blk["w_qkv"], blk["b_qkv"]).view(b, n, 3 * w)
            att = ops.attention(qkv[..., :w], qkv[..., w:2 * w], qkv[..., 2 * w:], heads=h, scale=64 ** -0.5, causal=True)
            x = ops.linear(att.view(b * n, w), blk["w_o"], blk["b_o"], residual=x)
            y = ops.layernorm(x, blk["ln2"][0], blk["ln2"][1], blk["ln2"][2])
            y = ops.linear(y, blk["w_fc"], blk["b_fc"], gelu=True)
            x = ops.linear(y, blk["w_pr"], blk["b_pr"], residual=x)
        z = ops.layernorm(x, *P["ln_final"])
        return z.view(b, n, w).to(self.model.positional_embedding.dtype)

    def forward(self, text):
        if torch.is_tensor(text):
            return self.encode_with_transformer(text)
        if self.tokenizer is None:
            raise RuntimeError("FrozenOpenCLIPEmbedder(B200): pass token ids ([B, 77] int64) or construct with tokenizer=open_clip.tokenize "
                               "(the BPE tokenizer is host string processing and is not part of this package)")
        return self.encode_with_transformer(self.tokenizer(text).to(self.model.positional_embedding.device))

    def encode(self, text):
        return self(text)


def video_to_uint8(video):
    """The tensor post-processing AFTER the path (app.py:90-94): [B, 3, T, H, W] in [-1, 1] -> uint8 [B, T, H, W, 3]."""
    return ops.video_to_uint8(video.contiguous())
