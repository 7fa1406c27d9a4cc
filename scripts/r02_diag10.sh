#!/bin/bash
# GroupNorm apply: tanh-form SiLU; occupancy / grid knobs.
mkdir -p gpurun_out/d10
python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "groupnorm or gn" -x 2>&1 | tail -3 > gpurun_out/d10/tests_gn.txt
for cfg in "4 4" "8 4" "4 5" "8 5" "6 4"; do set -- $cfg
  echo "== BPS=$1 MINB=$2" >> gpurun_out/d10/gn_bench.txt
  T2V_GN_BPS=$1 T2V_GN_MINB=$2 python scripts/gn_bench.py big 2>&1 | tail -8 >> gpurun_out/d10/gn_bench.txt
done
python scripts/gn_bench.py 2>&1 | tail -9 > gpurun_out/d10/gn_bench_small.txt
python -m pytest tests/test_model_gpu.py -q -m gpu -x 2>&1 | tail -5 > gpurun_out/d10/tests_model.txt
python bench.py > gpurun_out/d10/bench.json 2> gpurun_out/d10/bench.err
T2V_GN_BPS=8 python bench.py > gpurun_out/d10/bench_bps8.json 2>> gpurun_out/d10/bench.err
cat gpurun_out/d10/tests_gn.txt gpurun_out/d10/gn_bench.txt gpurun_out/d10/tests_model.txt gpurun_out/d10/bench.json gpurun_out/d10/bench_bps8.json
