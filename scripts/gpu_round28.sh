#!/bin/bash
C="lin320_res qkv320 geglu320 ff2_320 lin640_res qkv640 q640 geglu640 ff2_640 lin1280_res qkv1280 geglu1280 ff2_1280 conv320 conv640 conv1280 tconv320"
for bn in 0 128 160 256; do echo "== BN $bn"; BN=$bn timeout 300 python scripts/gemm_bench.py $C 2>&1 | grep "TF/s" | awk '{printf "%s %s | ", $1, $2} END {print ""}'; done
