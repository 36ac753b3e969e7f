"""SizeAdapter: top/left zero padding to a multiple of 64 and the matching crop
(reference size_adapter.py:10-52).  In the kernel path the crop is fused into
the estimator's store (``pds_subpixel_map(crop_top, crop_left)``), so
``PdsNetwork`` only asks this object for the pad amounts."""
from torch.nn import functional as F


class SizeAdapter(object):
    def __init__(self, minimum_size=64):
        self._minimum_size = minimum_size
        self._pixels_pad_to_width = None
        self._pixels_pad_to_height = None

    def padding_for(self, height, width):
        m = self._minimum_size
        return (-height) % m, (-width) % m

    def pad(self, network_input):
        height, width = network_input.size()[-2:]
        self._pixels_pad_to_height, self._pixels_pad_to_width = self.padding_for(height, width)
        return F.pad(network_input,
                     (self._pixels_pad_to_width, 0, self._pixels_pad_to_height, 0))

    def unpad(self, network_output):
        return network_output[..., self._pixels_pad_to_height:, self._pixels_pad_to_width:]
