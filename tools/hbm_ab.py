import os, sys, json
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests/golden")
import torch, bench
dev = torch.device("cuda")
u, r, ne, nr = bench.hbm_kernels(dev, 256, 60)
print("%s update %.2f us (%.0f GB/s) rot6d %.2f us (%.0f GB/s)" % (os.environ.get("REGEN_LIB_PATH", "default")[-12:], u * 1e3, 16 * ne / u / 1e6, r * 1e3, 60 * nr / r / 1e6))
