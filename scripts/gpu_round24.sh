#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "s2 or gaussian" 2>&1 | tail -n 3
timeout 300 python scripts/vae_encode_bench.py 2>&1 | tail -n 2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
