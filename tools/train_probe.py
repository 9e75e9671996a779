"""A few training steps of config 5's model (for ncu captures of the backward kernels):  python tools/train_probe.py [graphs] [precision] [steps]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                     # noqa: E402
import graphnets_b200 as gn      # noqa: E402
import workloads as W            # noqa: E402
import bench                     # noqa: E402

graphs = int(sys.argv[1]) if len(sys.argv) > 1 else 256
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
r = bench.train_leg(torch, gn, W, None, 1, graphs, steps=steps, precision=prec)
print(r["ms_per_step"], r["loss_first"], r["loss_last"])
