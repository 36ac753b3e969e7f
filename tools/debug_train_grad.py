import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import torch_port
from practicaldeepstereo_nips2018_b200 import PdsNetwork, loss as pds_loss, network_blocks
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(3)
left = torch.rand(1, 3, 128, 256, device='cuda') * 255
right = torch.rand(1, 3, 128, 256, device='cuda') * 255
gt = torch.rand(1, 128, 256, device='cuda') * 120
torch.manual_seed(0)
state = {k: v.clone() for k, v in PdsNetwork.default(127).state_dict().items()}
def step(kernels, dtype, batched=True, fused_loss=True, bench=False):
    torch.backends.cudnn.benchmark = bench
    network_blocks.USE_TRAINING_KERNELS = kernels
    net = PdsNetwork.default(127); net.load_state_dict(state); net = net.cuda().to(dtype).train()
    net._matching._batched_operation = batched
    cost = net(left.to(dtype), right.to(dtype))
    v = pds_loss.SubpixelCrossEntropy()(cost, gt) if (fused_loss and dtype == torch.float32) else torch_port.subpixel_cross_entropy(cost, gt.to(dtype), None, 1.0, 2)
    v.backward()
    return {k: p.grad.double() for k, p in net.named_parameters()}
d = step(False, torch.float64)
def dist(g):
    return (sum(float((g[k] - d[k]).norm()) ** 2 for k in d) / sum(float(d[k].norm()) ** 2 for k in d)) ** 0.5
print('ATen fp32                         %.3e' % dist(step(False, torch.float32, fused_loss=False)))
print('ATen fp32 + fused loss            %.3e' % dist(step(False, torch.float32)))
print('kernels, matching loop            %.3e' % dist(step(True, torch.float32, batched=False)))
print('kernels, batched matching         %.3e' % dist(step(True, torch.float32)))
print('ATen fp32 cudnn.benchmark         %.3e' % dist(step(False, torch.float32, fused_loss=False, bench=True)))
print('kernels batched cudnn.benchmark   %.3e' % dist(step(True, torch.float32, bench=True)))
