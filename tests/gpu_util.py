import numpy as np
import torch


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def tdict(params, device='cuda'):
    return {k: torch.from_numpy(v).to(device) for k, v in params.items()}


def max_abs(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)))) if a.size else 0.0


def load_module(module, params):
    module.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
    return module.cuda().eval()
