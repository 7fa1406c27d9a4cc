"""Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11; Random123) in numpy, and
the keep-mask `t2v_dropout_scale` must draw from it.  TEST INFRASTRUCTURE ONLY (tests/ and smoke may import it; the product
never does).

The generator is pinned against Random123's published known-answer vectors (tests/test_oracle.py); the kernel is then compared
bit for bit with `keep_mask` on the GPU (tests/test_lora_train_gpu.py).  The dropout itself restates nn.Dropout as the reference
uses it after lora_up (utils/lora.py:37,45-50) and in TemporalConvBlock (lvdm/modules/networks/openaimodel3d.py:280-296):
keep with probability 1 - p, scale kept values by 1 / (1 - p) — same distribution, not torch's random stream.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter, key):
    """counter: uint32 [..., 4], key: uint32 [..., 2] (broadcastable) -> uint32 [..., 4]."""
    c = [np.asarray(counter[..., i], dtype=np.uint32) for i in range(4)]
    k0 = np.asarray(key[..., 0], dtype=np.uint32)
    k1 = np.asarray(key[..., 1], dtype=np.uint32)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c[0].astype(np.uint64)
            p1 = M1 * c[2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK32).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK32).astype(np.uint32)
            c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
            k0 = (k0 + W0).astype(np.uint32)
            k1 = (k1 + W1).astype(np.uint32)
    return np.stack(c, axis=-1)


def keep_mask(n, keep_prob, seed, call_id):
    """The uint8 keep-mask of t2v_dropout_scale for n elements (n % 8 == 0): group g = i // 8 draws
    philox4x32_10((g_lo, g_hi, call_id, 0x74327662), (seed_lo, seed_hi)); element j of the group keeps iff its 16 random
    bits — low half of word j // 2 for even j, high half for odd j — are below round(keep_prob * 65536)."""
    assert n % 8 == 0
    g = np.arange(n // 8, dtype=np.uint64)
    ctr = np.stack([(g & MASK32).astype(np.uint32), (g >> np.uint64(32)).astype(np.uint32),
                    np.full(g.shape, call_id & 0xFFFFFFFF, dtype=np.uint32), np.full(g.shape, 0x74327662, dtype=np.uint32)], axis=-1)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    r = philox4x32_10(ctr, key[None, :])                                  # [groups, 4]
    halves = np.stack([r & np.uint32(0xFFFF), r >> np.uint32(16)], axis=-1).reshape(-1)   # [groups * 8]: lo0 hi0 lo1 hi1 ...
    thresh = np.uint32(int(np.rint(np.float32(keep_prob) * np.float32(65536.0))))
    return (halves < thresh).astype(np.uint8)
