"""Steps either side of the path: the OpenCLIP text tower (PARITY UNPINNED: open_clip is absent, the
checker is oracle/text_oracle.py) and the tensor post-processing of app.py:90-94."""
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
