"""SubpixelMap: sub-pixel MAP estimator (reference estimator.py:10-91) on one
sm_100a streaming kernel (csrc/estimator.cu)."""
import torch

from . import _capi


class SubpixelMap(object):
    """arg-max along disparity + soft-arg-max over a +-half_support_window
    neighbourhood.  Same constructor, validation and call signature as the
    reference; inference only; a plain object, not an nn.Module."""

    def __init__(self, half_support_window=4, disparity_step=2):
        if disparity_step < 1:
            raise ValueError('"disparity_step" should be positive integer.')
        if half_support_window < 1:
            raise ValueError('"half_support_window" should be positive integer.')
        if half_support_window % disparity_step != 0:
            raise ValueError('"half_support_window" should be multiple of the'
                             '"disparity_step"')
        self._disparity_step = disparity_step
        self._half_support_window = half_support_window

    def _differentiable(self, similarities, crop_top, crop_left, return_argmax):
        """Gradient-enabled calls (the reference's estimator is differentiable through the
        soft-arg-max, estimator.py:59-91): tensor expressions, any device, outside the inference
        hot path like every other gradient-enabled call of this package."""
        sim = similarities[..., crop_top:, crop_left:]
        depth = sim.size(1)
        radius = self._half_support_window // self._disparity_step
        best = torch.max(sim, dim=1, keepdim=True)[1]
        taps = best + torch.arange(-radius, radius + 1, device=sim.device).view(1, -1, 1, 1)
        inside = (taps >= 0) & (taps < depth)
        scores = torch.gather(sim, 1, taps.clamp(0, depth - 1))
        scores = torch.where(inside, scores, scores.new_full((), -float('inf')))
        weights = torch.softmax(scores, dim=1)
        tap_disparity = (taps * inside).to(sim.dtype) * self._disparity_step
        disparity = (weights * tap_disparity).sum(dim=1)
        return (disparity, best.squeeze(1)) if return_argmax else disparity

    def __call__(self, similarities, crop_top=0, crop_left=0, return_argmax=False):
        """similarities [B, D, H, W] (float32 or bfloat16, CUDA) -> disparity
        [B, H - crop_top, W - crop_left] float32.  The crop is SizeAdapter.unpad
        fused into the kernel's store."""
        if torch.is_grad_enabled() and similarities.requires_grad:
            return self._differentiable(similarities, crop_top, crop_left, return_argmax)
        _capi.require_cuda(similarities)
        if similarities.dim() != 4:
            raise ValueError('similarities should have indices [batch, disparity, y, x]')
        if similarities.dtype == torch.float64:
            raise TypeError('pds_b200 SubpixelMap supports float32 / bfloat16 cost volumes')
        sim = similarities.detach().contiguous()
        B, D, H, W = sim.shape
        out = torch.empty((B, H - crop_top, W - crop_left), dtype=torch.float32,
                          device=sim.device)
        idx = torch.empty_like(out, dtype=torch.int64) if return_argmax else None
        with torch.cuda.device(sim.device):
            _capi.check(_capi.lib().pds_subpixel_map(
                _capi.ptr(sim), _capi.ptr(out), _capi.ptr(idx) if idx is not None else None,
                B, D, H, W, self._half_support_window, self._disparity_step, crop_top,
                crop_left, _capi.dtype_code(sim), _capi.stream_ptr(sim.device)))
        return (out, idx) if return_argmax else out
