"""Runs the VAE decode at the bench shape so ncu can list its kernels (numbers under ncu are not reported)."""
import sys
sys.path.insert(0, ".")
import torch
import bench
print(bench.bench_vae_decode(torch.device("cuda", 0)))
