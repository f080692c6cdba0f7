"""Restatement of the third-party package ``alias_free_torch~=0.0.6``.

TEST INFRASTRUCTURE (oracle).  The reference pins this dependency in its
requirements.txt but does not vendor it; BigVGAN imports it at
TrainingInterfaces/Spectrogram_to_Wave/BigVGAN/AMP.py:8-9 (``from
alias_free_torch import *`` -- so this module must re-export ``torch``, ``nn``,
``F`` and ``math`` as the original does) and
InferenceInterfaces/InferenceArchitectures/InferenceBigVGAN.py:8.

Published algorithm (github.com/junjun3518/alias-free-torch, v0.0.6):
  * ``kaiser_sinc_filter1d(cutoff, half_width, kernel_size)``: Kaiser-windowed
    sinc low-pass, normalised to unit sum.
  * ``UpSample1d(ratio=2, kernel_size=12)``: replicate-pad ``k/ratio-1`` both
    sides, ``ratio * conv_transpose1d(stride=ratio, depthwise)``, crop.
  * ``LowPassFilter1d`` / ``DownSample1d(ratio=2, kernel_size=12)``:
    replicate-pad ``(k/2-1, k/2)``, depthwise ``conv1d(stride=ratio)``.
  * ``Activation1d``: upsample -> activation -> downsample.

"Parity unpinned": the reference has no test touching this boundary.  The
filter taps are cross-checked in tests/test_oracle.py against the values probed
in SURVEY.md section 8c and against transformers' independent copy.
"""
import math

import torch
from torch import nn
from torch.nn import functional as F

__all__ = ["torch", "nn", "F", "math", "sinc", "kaiser_sinc_filter1d", "LowPassFilter1d",
           "UpSample1d", "DownSample1d", "Activation1d"]


def sinc(x):
    return torch.where(x == 0, torch.tensor(1.0, device=x.device, dtype=x.dtype),
                       torch.sin(math.pi * x) / math.pi / x)


def kaiser_sinc_filter1d(cutoff, half_width, kernel_size):
    even = kernel_size % 2 == 0
    half_size = kernel_size // 2
    delta_f = 4 * half_width
    attenuation = 2.285 * (half_size - 1) * math.pi * delta_f + 7.95
    if attenuation > 50.0:
        beta = 0.1102 * (attenuation - 8.7)
    elif attenuation >= 21.0:
        beta = 0.5842 * (attenuation - 21) ** 0.4 + 0.07886 * (attenuation - 21.0)
    else:
        beta = 0.0
    window = torch.kaiser_window(kernel_size, beta=beta, periodic=False)
    if even:
        time = torch.arange(-half_size, half_size) + 0.5
    else:
        time = torch.arange(kernel_size) - half_size
    if cutoff == 0:
        filt = torch.zeros_like(time)
    else:
        filt = 2 * cutoff * window * sinc(2 * cutoff * time)
        filt = filt / filt.sum()
    return filt.view(1, 1, kernel_size)


class LowPassFilter1d(nn.Module):
    def __init__(self, cutoff=0.5, half_width=0.6, stride=1, padding=True, padding_mode="replicate",
                 kernel_size=12):
        super().__init__()
        if cutoff < -0.0:
            raise ValueError("Minimum cutoff must be larger than zero.")
        if cutoff > 0.5:
            raise ValueError("A cutoff above 0.5 does not make sense.")
        self.kernel_size = kernel_size
        self.even = kernel_size % 2 == 0
        self.pad_left = kernel_size // 2 - int(self.even)
        self.pad_right = kernel_size // 2
        self.stride = stride
        self.padding = padding
        self.padding_mode = padding_mode
        self.register_buffer("filter", kaiser_sinc_filter1d(cutoff, half_width, kernel_size))

    def forward(self, x):
        _, channels, _ = x.shape
        if self.padding:
            x = F.pad(x, (self.pad_left, self.pad_right), mode=self.padding_mode)
        return F.conv1d(x, self.filter.expand(channels, -1, -1), stride=self.stride, groups=channels)


class UpSample1d(nn.Module):
    def __init__(self, ratio=2, kernel_size=None):
        super().__init__()
        self.ratio = ratio
        self.kernel_size = int(6 * ratio // 2) * 2 if kernel_size is None else kernel_size
        self.stride = ratio
        self.pad = self.kernel_size // ratio - 1
        self.pad_left = self.pad * self.stride + (self.kernel_size - self.stride) // 2
        self.pad_right = self.pad * self.stride + (self.kernel_size - self.stride + 1) // 2
        self.register_buffer("filter", kaiser_sinc_filter1d(cutoff=0.5 / ratio, half_width=0.6 / ratio,
                                                            kernel_size=self.kernel_size))

    def forward(self, x):
        _, channels, _ = x.shape
        x = F.pad(x, (self.pad, self.pad), mode="replicate")
        x = self.ratio * F.conv_transpose1d(x, self.filter.expand(channels, -1, -1), stride=self.stride,
                                            groups=channels)
        return x[..., self.pad_left:-self.pad_right]


class DownSample1d(nn.Module):
    def __init__(self, ratio=2, kernel_size=None):
        super().__init__()
        self.ratio = ratio
        self.kernel_size = int(6 * ratio // 2) * 2 if kernel_size is None else kernel_size
        self.lowpass = LowPassFilter1d(cutoff=0.5 / ratio, half_width=0.6 / ratio, stride=ratio,
                                       kernel_size=self.kernel_size)

    def forward(self, x):
        return self.lowpass(x)


class Activation1d(nn.Module):
    def __init__(self, activation, up_ratio=2, down_ratio=2, up_kernel_size=12, down_kernel_size=12):
        super().__init__()
        self.up_ratio = up_ratio
        self.down_ratio = down_ratio
        self.act = activation
        self.upsample = UpSample1d(up_ratio, up_kernel_size)
        self.downsample = DownSample1d(down_ratio, down_kernel_size)

    def forward(self, x):
        x = self.upsample(x)
        x = self.act(x)
        x = self.downsample(x)
        return x
