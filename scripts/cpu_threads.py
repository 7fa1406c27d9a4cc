"""How does the oracle CPU sample scale with threads on this box? (sizing the cpu_baseline sample)"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
for thr in (16, 32, 64):
    bench_run, tf, desc = bench.cpu_sample(thr, t_frames=1)
    t = bench_run()
    print(f"threads={thr} frames=1: {t:.1f} s", flush=True)
