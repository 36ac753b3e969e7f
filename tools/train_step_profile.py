"""Kernel table of one PdsNetwork training step (torch.profiler, CUDA time by kernel name)."""
import os, sys
import torch
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from practicaldeepstereo_nips2018_b200 import PdsNetwork, loss as pds_loss
torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
H, W, md = 540, 960, 255
torch.manual_seed(0)
net = PdsNetwork.default(md).cuda().train()
opt = torch.optim.RMSprop(net.parameters(), lr=1e-2)
crit = pds_loss.SubpixelCrossEntropy()
left = torch.rand(1, 3, H, W, device='cuda') * 255
right = torch.rand(1, 3, H, W, device='cuda') * 255
gt = torch.rand(1, H, W, device='cuda') * (md - 1)
def one():
    opt.zero_grad()
    v = crit(net(left, right), gt)
    v.backward()
    opt.step()
for _ in range(2):
    one()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    one()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=40, max_name_column_width=90))
