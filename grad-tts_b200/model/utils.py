"""Mask / length helpers feeding shapes to the hot path (reference model/utils.py:6-39). Stay in PyTorch."""
import torch


def sequence_mask(length, max_length=None):
    # model/utils.py:6-10
    if max_length is None:
        max_length = length.max()
    x = torch.arange(int(max_length), dtype=length.dtype, device=length.device)
    return x.unsqueeze(0) < length.unsqueeze(1)


def fix_len_compatibility(length, num_downsamplings_in_unet=2):
    # model/utils.py:13-17: round up to a multiple of 2**n
    q = 2 ** num_downsamplings_in_unet
    return ((int(length) + q - 1) // q) * q


def generate_path(duration, mask):
    # model/utils.py:26-39: hard monotonic alignment from integer durations
    b, t_x, t_y = mask.shape
    cum = torch.cumsum(duration, 1)
    path = sequence_mask(cum.reshape(b * t_x), t_y).to(mask.dtype).view(b, t_x, t_y)
    path = path - torch.nn.functional.pad(path, (0, 0, 1, 0, 0, 0))[:, :-1]
    return path * mask


def duration_loss(logw, logw_, lengths):
    # model/utils.py:42-44
    return torch.sum((logw - logw_) ** 2) / torch.sum(lengths)
