import torch
x = torch.zeros(1, device="cuda")
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3): x.add_(1)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
N = 2000
with torch.cuda.graph(g):
    for _ in range(N): x.add_(1)
for _ in range(3): g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
print("graph: dependent tiny-kernel chain:", e0.elapsed_time(e1) * 1e3 / N, "us per node")
e0.record()
for _ in range(N): x.add_(1)
e1.record(); torch.cuda.synchronize()
print("eager stream launches:", e0.elapsed_time(e1) * 1e3 / N, "us per kernel")
