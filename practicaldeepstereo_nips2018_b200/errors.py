"""Disparity error metrics (reference errors.py:9-74): same functions, same return values.
CUDA float32 inputs go through one fused kernel (csrc/errors.cu) that produces the pixel-wise
maps and the sums of both metrics in a single pass; everything else (CPU tensors, other dtypes,
the median variant) runs the reference's tensor expressions."""
import torch

from . import _capi


def _fused(estimated_disparity, ground_truth_disparity):
    return (torch.is_tensor(estimated_disparity) and torch.is_tensor(ground_truth_disparity)
            and estimated_disparity.is_cuda and ground_truth_disparity.is_cuda
            and estimated_disparity.dtype == torch.float32
            and ground_truth_disparity.dtype == torch.float32
            and estimated_disparity.shape == ground_truth_disparity.shape
            and not estimated_disparity.requires_grad)


def compute_errors(estimated_disparity, ground_truth_disparity, n=3.0):
    """Both metrics from one pass: (pixelwise_absolute_error, mean_absolute_error,
    pixelwise_n_pixels_error, percentage_of_pixels_with_error)."""
    if not _fused(estimated_disparity, ground_truth_disparity):
        pixelwise_abs, mean_abs = compute_absolute_error(estimated_disparity, ground_truth_disparity)
        pixelwise_bad, bad = compute_n_pixels_error(estimated_disparity, ground_truth_disparity, n)
        return pixelwise_abs, mean_abs, pixelwise_bad, bad
    est, gt = estimated_disparity.contiguous(), ground_truth_disparity.contiguous()
    pixelwise_abs, pixelwise_bad = torch.empty_like(est), torch.empty_like(est)
    sums = torch.empty(3, dtype=torch.float64, device=est.device)
    with torch.cuda.device(est.device):
        _capi.check(_capi.lib().pds_disparity_errors(
            _capi.ptr(est), _capi.ptr(gt), _capi.ptr(pixelwise_abs), _capi.ptr(pixelwise_bad),
            est.numel(), float(n), _capi.ptr(sums), _capi.stream_ptr(est.device)))
    total, known, bad = sums.tolist()                    # the reference's .item() synchronisation
    if known == 0:
        return pixelwise_abs, 0.0, pixelwise_bad, 0.0
    return pixelwise_abs, total / known, pixelwise_bad, bad / known * 100


def compute_absolute_error(estimated_disparity, ground_truth_disparity, use_mean=True):
    """Pixel-wise and mean (or median) absolute error; unknown ground truth (inf) is skipped
    and shown as zero; 0 when nothing is known (errors.py:9-41)."""
    if use_mean and _fused(estimated_disparity, ground_truth_disparity):
        pixelwise_abs, mean_abs, _, _ = compute_errors(estimated_disparity, ground_truth_disparity)
        return pixelwise_abs, mean_abs
    absolute_difference = (estimated_disparity - ground_truth_disparity).abs()
    unknown = torch.isinf(ground_truth_disparity)
    pixelwise = absolute_difference.clone()
    pixelwise[unknown] = 0
    known = absolute_difference[~unknown]
    if known.numel() == 0:
        return pixelwise, 0.0
    return pixelwise, (known.mean() if use_mean else known.median()).item()


def compute_n_pixels_error(estimated_disparity, ground_truth_disparity, n=3.0):
    """Pixel-wise n-pixels error and the percentage of known pixels whose absolute error
    exceeds n (errors.py:44-74)."""
    if _fused(estimated_disparity, ground_truth_disparity):
        _, _, pixelwise_bad, bad = compute_errors(estimated_disparity, ground_truth_disparity, n)
        return pixelwise_bad, bad
    unknown = torch.isinf(ground_truth_disparity)
    exceeds = (estimated_disparity - ground_truth_disparity).abs().gt(n).float()
    pixelwise = exceeds.clone()
    pixelwise[unknown] = 0.0
    known = exceeds[~unknown]
    if known.numel() == 0:
        return pixelwise, 0.0
    return pixelwise, known.mean().item() * 100
