#!/bin/bash
C="lin320_res lin320 qkv320 geglu320 ff2_320 lin640_res geglu640 tconv320 lin1280_res"
for t in 0 0x040000 0x0a0000; do echo "== tune $t"; T2V_GEMM_TUNE=$t timeout 300 python scripts/gemm_bench.py $C 2>&1 | tail -n 9; done
