"""One pass of the ScSPM pipeline of bench.py (for an ncu launch list)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
rate, _ = bench.scspm_images_per_s(torch.device("cuda:0"), reps=1)
print("images/s", rate)
