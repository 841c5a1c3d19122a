"""Drop-in for reference ``gantts/seqloss.py`` (sequence_mask :9-20, MaskedMSELoss :27-43) on CUDA."""
import torch
from torch import nn

from . import ops


def sequence_mask(sequence_length, max_len=None):
    """(B,) lengths -> (B, max_len) float {0,1} mask; reference gantts/seqloss.py:9-20."""
    if not torch.is_tensor(sequence_length):
        raise RuntimeError("sequence_mask expects a tensor of lengths")
    if max_len is None:
        max_len = int(sequence_length.max().item())     # same host read as `.data.max()` upstream
    return ops.sequence_mask(sequence_length, int(max_len))


class MaskedMSELoss(nn.Module):
    """sum(((input - target) * mask)^2) / sum(mask) with a (B, T, 1) mask -- i.e. normalised by
    the number of valid FRAMES (reference gantts/seqloss.py:27-43)."""

    def __init__(self):
        super(MaskedMSELoss, self).__init__()

    def forward(self, input, target, lengths=None, mask=None, max_len=None):
        if lengths is None and mask is None:
            raise RuntimeError("Should provide either lengths or mask")
        if mask is None:
            if not lengths.is_cuda:
                lengths = lengths.to(input.device)
            mask = sequence_mask(lengths, max_len).unsqueeze(-1)
        if mask.shape[:2] != input.shape[:2]:
            raise RuntimeError("mask %s does not match input %s" % (tuple(mask.shape), tuple(input.shape)))
        return ops.masked_mse(input, target, mask)
