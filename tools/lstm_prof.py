"""One forward+backward of a cfg3-width BiLSTM layer stack (for ncu captures of the recurrence kernels)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__
__graft_entry__.build()
import gantts_b200
dev = torch.device("cuda:0")
torch.manual_seed(0)
B, T = int(os.environ.get("B", 16)), int(os.environ.get("T", 500))
m = gantts_b200.models.LSTMRNN(177, 177, 1, 512, bidirectional=True).to(dev).train()
x = torch.randn(B, T, 177, device=dev, requires_grad=True)
for _ in range(2):
    y = m(x, [T] * B)
    y.backward(torch.ones_like(y))
torch.cuda.synchronize()
print("ok")
