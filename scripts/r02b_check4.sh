#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/attn_ablate.py "" p0 p2 p3 p4 p5 p6 p8 q2 q3 q4 q5 q6 q8 v1 2>&1 | tee gpurun_out/b4_attn_ablate.txt | tail -n 16
