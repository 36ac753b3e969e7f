"""SubpixelCrossEntropy (reference loss.py:16-78): same constructor, same forward signature, same
value and gradients.  CUDA float32 tensors go through two fused kernels (csrc/loss.cu: one pass
over the similarity volume forward, one backward) behind a custom autograd Function; everything
else (CPU tensors, other dtypes) evaluates the reference's formula with tensor expressions."""
import torch
from torch import nn

from . import _capi


def _none_ptr(t):
    return _capi.ptr(t) if t is not None else None


class _FusedSubpixelCrossEntropy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, similarities, ground_truth, weights, diversity, disparity_step):
        sim = similarities.contiguous()
        gt = ground_truth.contiguous().float()
        w = weights.contiguous().float() if weights is not None else None
        B, D, H, W = sim.shape
        entropy = torch.empty((B, H, W), dtype=torch.float32, device=sim.device)
        lse, sum_pt = torch.empty_like(entropy), torch.empty_like(entropy)
        sums = torch.empty(2, dtype=torch.float64, device=sim.device)
        with torch.cuda.device(sim.device):
            _capi.check(_capi.lib().pds_subpixel_cross_entropy_forward(
                _capi.ptr(sim), _capi.ptr(gt), _none_ptr(w), _capi.ptr(entropy), _capi.ptr(lse),
                _capi.ptr(sum_pt), _capi.ptr(sums), B, D, H, W, float(diversity), int(disparity_step),
                _capi.stream_ptr(sim.device)))
        ctx.save_for_backward(sim, gt, w if w is not None else sim.new_empty(0), entropy, lse, sum_pt, sums)
        ctx.has_weights = w is not None
        ctx.args = (float(diversity), int(disparity_step))
        denominator = sums[1] + 1e-15 if w is not None else sums[1]      # loss.py:74-78
        return (sums[0] / denominator).float()

    @staticmethod
    def backward(ctx, upstream):
        sim, gt, w, entropy, lse, sum_pt, sums = ctx.saved_tensors
        w = w if ctx.has_weights else None
        B, D, H, W = sim.shape
        up = upstream.reshape(1).float().contiguous()
        grad_sim = torch.empty_like(sim) if ctx.needs_input_grad[0] else None
        want_w = ctx.has_weights and ctx.needs_input_grad[2]
        grad_w = torch.empty_like(entropy) if want_w else None
        if grad_sim is None and grad_w is None:
            return None, None, None, None, None
        scratch = grad_sim if grad_sim is not None else torch.empty_like(sim)
        with torch.cuda.device(sim.device):
            _capi.check(_capi.lib().pds_subpixel_cross_entropy_backward(
                _capi.ptr(sim), _capi.ptr(gt), _none_ptr(w), _capi.ptr(entropy), _capi.ptr(lse),
                _capi.ptr(sum_pt), _capi.ptr(sums), _capi.ptr(up), _capi.ptr(scratch), _none_ptr(grad_w),
                B, D, H, W, ctx.args[0], ctx.args[1], _capi.stream_ptr(sim.device)))
        return grad_sim, None, grad_w, None, None


class SubpixelCrossEntropy(nn.Module):
    def __init__(self, diversity=1.0, disparity_step=2):
        """diversity: scale of the target Laplace distribution centred at the sub-pixel ground
        truth; disparity_step: disparity difference between neighbouring indices of `similarities`."""
        super().__init__()
        self._diversity = diversity
        self._disparity_step = disparity_step

    def _tensor_expression(self, similarities, ground_truth_disparities, weights):
        # loss.py:52-78 without the Python loop over the disparity axis
        known = ~torch.isinf(ground_truth_disparities.detach())
        log_p = torch.log_softmax(similarities, dim=1)
        disparities = torch.arange(similarities.size(1), device=similarities.device,
                                   dtype=similarities.dtype).view(1, -1, 1, 1) * self._disparity_step
        p_target = torch.exp(-(ground_truth_disparities.unsqueeze(1) - disparities).abs() / self._diversity) \
            / (2 * self._diversity)
        entropy = -(log_p * p_target).sum(dim=1)[known] / p_target.sum(dim=1)[known]
        if weights is not None:
            w = weights[known]
            return (w * entropy).sum() / (w.sum() + 1e-15)
        return entropy.mean()

    def forward(self, similarities, ground_truth_disparities, weights=None):
        """similarities [B, D, H, W], ground_truth_disparities [B, H, W] (inf = unknown),
        weights [B, H, W] or None -> scalar loss."""
        if (similarities.is_cuda and similarities.dtype == torch.float32 and similarities.dim() == 4
                and tuple(ground_truth_disparities.shape) == (similarities.size(0),) + tuple(similarities.shape[2:])
                and (weights is None or weights.shape == ground_truth_disparities.shape)):
            return _FusedSubpixelCrossEntropy.apply(similarities, ground_truth_disparities, weights,
                                                    self._diversity, self._disparity_step)
        return self._tensor_expression(similarities, ground_truth_disparities, weights)
