import torch.nn.functional as F


def memory_efficient_attention(q, k, v, attn_bias=None, op=None):
    """xformers layout [B, M, H, K] -> torch SDPA (flash / mem-efficient kernels) -> [B, M, H, K]."""
    assert attn_bias is None
    out = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
    return out.transpose(1, 2)
