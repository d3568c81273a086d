"""VirtualBatchNorm with the reference's interface (estorch/modules.py:6-58).

Same constructor (``num_features``, ``eps=1e-5``), same learnable ``weight`` /
``bias`` per channel, same two-call protocol: the first ``forward`` after a
reset takes the *reference batch*, stores per-(C,H,W) mean and unbiased
variance over the batch dimension and normalises it; the second ``forward``
normalises the real batch with those statistics and clears them
(modules.py:48-58).  Statistics are plain attributes, not buffers, so they are
not part of ``state_dict`` -- as in the reference.
"""
import torch
from torch import nn


class VirtualBatchNorm(nn.Module):
    def __init__(self, num_features, eps=1e-5):
        super().__init__()
        self.num_features = num_features
        self.eps = eps
        self.mean = None
        self.var = None
        self.weight = nn.Parameter(torch.ones(num_features))
        self.bias = nn.Parameter(torch.zeros(num_features))

    def reset_parameters(self):
        nn.init.ones_(self.weight)
        nn.init.zeros_(self.bias)

    def normalize(self, x):
        shape = (1, self.num_features, 1, 1)
        scale = self.weight.view(shape)
        shift = self.bias.view(shape)
        return (x - self.mean) / torch.sqrt(self.var + self.eps) * scale + shift

    def forward(self, x):
        first_call = self.mean is None and self.var is None
        if first_call:
            self.mean = x.mean(dim=0, keepdim=True)
            self.var = x.var(dim=0, keepdim=True)   # unbiased, like torch.var's default
        out = self.normalize(x)
        if not first_call:
            self.mean = None
            self.var = None
        return out
