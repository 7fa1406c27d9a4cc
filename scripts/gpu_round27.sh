#!/bin/bash
for o in 2 4 6 8 12; do echo "== occ $o"; T2V_GN_OCC=$o timeout 200 python scripts/gn_bench.py 2>&1 | grep -E "hw=2560 c=320|hw=640 c=640|hw=640 c=1280|hw=160 c=1280 " | sed 's/, cluster.*//'; done
