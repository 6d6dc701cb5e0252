import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/rayuela.jl_b200')
from rayuela_b200 import core
dev = torch.device('cuda'); n, d, m = 1000000, 128, 8
X = torch.randn(n, d, device=dev); B = torch.randint(0, 256, (n, m), device=dev, dtype=torch.uint8)
for _ in range(3): core.fast_bin_matmul(X, B)
torch.cuda.synchronize()
