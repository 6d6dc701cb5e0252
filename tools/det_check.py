import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/rayuela.jl_b200')
import bench
dev = torch.device('cuda')
def chk(t): return int(t.double().sum().item()*1e3) , float(t.double().abs().sum().item())
for rep in range(2):
    Xt, _ = bench.make_data(50000, 1, 128, seed=4000, device=dev, rank=0)
    C = bench.train_codebooks(Xt, 16, dev)
    X0 = bench.make_data(125000, 1, 128, seed=5000, device=dev, rank=0)[0]
    B0 = torch.randint(0, 256, (125000, 16), device=dev, dtype=torch.uint8, generator=torch.Generator(device=dev).manual_seed(6000))
    print(rep, 'Xt', chk(Xt), 'C', chk(C), 'X0', chk(X0), 'B0', int(B0.long().sum()))
