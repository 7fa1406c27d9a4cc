"""Stand-in for the `xformers` package (not installed here), used ONLY by scripts/ref_gpu_bench.py to exercise the
unmodified reference's own flash path: with this on sys.path `lvdm/modules/attention.py:7-13` sets XFORMERS_IS_AVAILBLE
and spatial CrossAttention runs `efficient_forward` (:166-240), whose single library call is routed to torch SDPA."""
from . import ops  # noqa: F401
