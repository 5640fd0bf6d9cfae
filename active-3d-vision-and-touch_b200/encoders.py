"""Vertex-feature encoders of the deformation network with the reference's constructor arguments, sub-module
layout and state-dict names (reference checkpoints load unchanged):

    Positional_Encoder(input_size)   pterotactyl/reconstruction/vision/model.py:367-399
    Mask_Encoder(input_size)         pterotactyl/reconstruction/vision/model.py:402-414

The NeRF embedding + concatenation (20 sin/cos, 20 multiplies, 2 cat per call in the reference) is one kernel
(ops.nerf_embed -> ptk_nerf_embed_fwd/bwd); the three small Linear layers stay torch modules (library GEMMs,
outside the scope of the hand-written path -- DESIGN.md section 7).
"""
import torch.nn as nn

from . import ops


class Positional_Encoder(nn.Module):
    def __init__(self, input_size):
        super(Positional_Encoder, self).__init__()
        self.model = nn.Sequential(
            nn.Linear(63, input_size // 4),  # 10 NeRF frequencies x (sin, cos) x 3 + the positions
            nn.ReLU(inplace=True),
            nn.Linear(input_size // 4, input_size // 2),
            nn.ReLU(inplace=True),
            nn.Linear(input_size // 2, input_size),
        )

    def nerf_embedding(self, points):
        """The 60-wide sin/cos part, as the reference's method returns it (model.py:381-391)."""
        return ops.nerf_embed(points)[..., :60]

    def forward(self, positions):
        shape = positions.shape
        x = ops.nerf_embed(positions.contiguous().view(shape[0] * shape[1], -1))
        return self.model(x).view(shape[0], shape[1], -1)


class Mask_Encoder(nn.Module):
    def __init__(self, input_size):
        super(Mask_Encoder, self).__init__()
        self.model = nn.Sequential(nn.Embedding(4, input_size))

    def forward(self, mask):
        shape = mask.shape
        return self.model(mask.contiguous().view(-1, 1).long()).view(shape[0], shape[1], -1)
