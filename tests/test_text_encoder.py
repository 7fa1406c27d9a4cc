"""Steps either side of the path: the OpenCLIP text tower (open_clip is absent; the checker, oracle/text_oracle.py,
is pinned on the independent implementation of the same tower that IS installed: transformers.CLIPTextModel) and the tensor
post-processing of app.py:90-94."""
import pytest
import torch


def test_text_tower_census_vit_h_14():
    """ViT-H-14 text tower: 24 blocks, width 1024 — 354 M parameters with open_clip's key names."""
    from t2v_turbo_b200.text_encoder import FrozenOpenCLIPEmbedder
    with torch.device("meta"):
        m = FrozenOpenCLIPEmbedder(layer="penultimate")
    sd = m.state_dict()
    assert "model.transformer.resblocks.23.attn.in_proj_weight" in sd and "model.positional_embedding" in sd and "model.ln_final.bias" in sd
    assert sum(v.numel() for v in sd.values()) == 354_032_641 and m.layer_idx == 1


def openclip_to_hf_clip(sd, layers):
    """open_clip text-tower keys -> transformers CLIPTextModel keys (the fused in_proj is q | k | v)."""
    out = {"text_model.embeddings.token_embedding.weight": sd["model.token_embedding.weight"],
           "text_model.embeddings.position_embedding.weight": sd["model.positional_embedding"],
           "text_model.final_layer_norm.weight": sd["model.ln_final.weight"], "text_model.final_layer_norm.bias": sd["model.ln_final.bias"]}
    for i in range(layers):
        a, b = f"model.transformer.resblocks.{i}.", f"text_model.encoder.layers.{i}."
        for j, n in enumerate(("q_proj", "k_proj", "v_proj")):
            out[b + f"self_attn.{n}.weight"] = sd[a + "attn.in_proj_weight"].chunk(3, 0)[j]
            out[b + f"self_attn.{n}.bias"] = sd[a + "attn.in_proj_bias"].chunk(3, 0)[j]
        for src, dst in (("attn.out_proj", "self_attn.out_proj"), ("ln_1", "layer_norm1"), ("ln_2", "layer_norm2"), ("mlp.c_fc", "mlp.fc1"),
                         ("mlp.c_proj", "mlp.fc2")):
            out[b + dst + ".weight"], out[b + dst + ".bias"] = sd[a + src + ".weight"], sd[a + src + ".bias"]
    return out


def test_text_oracle_pinned_on_transformers_clip_text_model():
    """open_clip is absent, but `transformers` ships an independent implementation of the SAME text tower (CLIPTextModel: the MS
    pipeline's own text encoder, pipeline/t2v_turbo_ms_pipeline.py:36-44, and — with hidden_act="gelu" — OpenCLIP ViT-H's): the
    oracle restatement must agree with it on shared weights, for the last layer and for FrozenOpenCLIPEmbedder's penultimate
    layer (condition.py:257-283: skip the last block, then ln_final) = hidden_states[-2] -> final_layer_norm."""
    transformers = pytest.importorskip("transformers")
    from oracle.text_oracle import text_forward
    from oracle.weights import seeded_state_dict
    from t2v_turbo_b200.text_encoder import FrozenOpenCLIPEmbedder
    w, layers, heads, vocab = 256, 4, 4, 1000
    m = FrozenOpenCLIPEmbedder(layer="penultimate", width=w, layers=layers, heads=heads, vocab_size=vocab)
    sd = seeded_state_dict(m.state_dict(), 41)
    sd["model.positional_embedding"] = sd["model.positional_embedding"] * 0.3
    cfg = transformers.CLIPTextConfig(vocab_size=vocab, hidden_size=w, intermediate_size=4 * w, num_hidden_layers=layers, num_attention_heads=heads,
                                      max_position_embeddings=77, hidden_act="gelu", layer_norm_eps=1e-5, eos_token_id=vocab - 1)
    hf = transformers.CLIPTextModel(cfg).eval()
    missing, unexpected = hf.load_state_dict(openclip_to_hf_clip(sd, layers), strict=False)
    assert not unexpected and all("position_ids" in k for k in missing), (missing, unexpected)
    tokens = torch.randint(0, vocab, (3, 77), generator=torch.Generator().manual_seed(42))
    with torch.no_grad():
        out = hf(input_ids=tokens, output_hidden_states=True)
        hf_last = out.last_hidden_state
        hf_pen = hf.text_model.final_layer_norm(out.hidden_states[-2])
        ours_last, ours_pen = text_forward(sd, tokens, heads=heads, layer_idx=0), text_forward(sd, tokens, heads=heads, layer_idx=1)
    rel = lambda a, b: ((a - b).norm() / b.norm()).item()      # noqa: E731
    assert rel(ours_last, hf_last) < 1e-5 and rel(ours_pen, hf_pen) < 1e-5, (rel(ours_last, hf_last), rel(ours_pen, hf_pen))
    assert rel(ours_pen, hf_last) > 0.1                     # the two layers are genuinely different outputs


@pytest.mark.gpu
def test_text_tower_vs_oracle(cuda_device):
    from oracle.text_oracle import text_forward
    from oracle.weights import seeded_state_dict
    from t2v_turbo_b200.text_encoder import FrozenOpenCLIPEmbedder
    m = FrozenOpenCLIPEmbedder(layer="penultimate", width=256, layers=4, heads=4, vocab_size=1000)
    sd = seeded_state_dict(m.state_dict(), 41)
    sd["model.positional_embedding"] = sd["model.positional_embedding"] * 0.3
    m.load_state_dict(sd)
    m = m.cuda()
    tokens = torch.randint(0, 1000, (3, 77), generator=torch.Generator().manual_seed(42))
    with torch.no_grad():
        ref = text_forward(sd, tokens, heads=4, layer_idx=1)
        ref_last = text_forward(sd, tokens, heads=4, layer_idx=0)
    z = m(tokens.cuda())
    rel = ((z.float().cpu() - ref).norm() / ref.norm()).item()
    print(f"\n[text tower penultimate] rel-L2 vs oracle {rel:.3e} (last-layer output differs by {((ref_last - ref).norm() / ref.norm()).item():.2f})")
    assert z.shape == (3, 77, 256) and rel <= 2.0e-2
    # causality: changing a later token must not change earlier positions
    t2 = tokens.clone()
    t2[:, 40:] = (t2[:, 40:] + 7) % 1000
    z2 = m(t2.cuda())
    assert torch.equal(z[:, :40], z2[:, :40]) and not torch.equal(z[:, 40:], z2[:, 40:])
    with pytest.raises(RuntimeError):
        m(["a prompt"])


@pytest.mark.gpu
def test_video_to_uint8_matches_app_postprocess(cuda_device):
    from t2v_turbo_b200.text_encoder import video_to_uint8
    v = (torch.randn(2, 3, 4, 16, 24, generator=torch.Generator().manual_seed(43)) * 0.8).bfloat16()
    ref = []
    for vid in v:   # app.py:90-94
        x = torch.clamp(vid.float(), -1.0, 1.0).permute(1, 0, 2, 3)
        x = (x + 1.0) / 2.0
        ref.append((x * 255).to(torch.uint8).permute(0, 2, 3, 1))
    out = video_to_uint8(v.cuda())
    assert out.dtype == torch.uint8 and torch.equal(out.cpu(), torch.stack(ref))
