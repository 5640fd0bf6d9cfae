"""Vertex-feature encoders of the deformation network with the reference's constructor arguments, sub-module
layout and state-dict names (reference checkpoints load unchanged):

    Positional_Encoder(input_size)   pterotactyl/reconstruction/vision/model.py:367-399
    Mask_Encoder(input_size)         pterotactyl/reconstruction/vision/model.py:402-414

`Positional_Encoder.forward` is ONE launch (ops.vertex_front -> ptk_vertex_front_fwd): NeRF embedding (20 sin/cos,
20 multiplies, 2 cat per call in the reference), the three Linear layers and their ReLUs with the activations held in
shared memory.  `vertex_features(...)` is the whole front of a deformation iteration -- positional MLP + mask-token
embedding row + optional pooled image features (model.py:229-236, 261-267, 274-279) -- in that same single launch,
writing the GCN's layer-0 input directly; `recon.ChartDeformer` callers use it, the reference's unedited
Deformation.forward reaches the fused MLP through the patched Positional_Encoder class and keeps its own adds.
Widths beyond the kernel's shared-memory budget (input_size > 448) take the embedding kernel + torch Linear layers.
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
        self.fused = ops.vertex_front_supported(input_size)

    def nerf_embedding(self, points):
        """The 60-wide sin/cos part, as the reference's method returns it (model.py:381-391)."""
        return ops.nerf_embed(points)[..., :60]

    def _weights(self):
        l1, l2, l3 = self.model[0], self.model[2], self.model[4]
        return l1.weight, l1.bias, l2.weight, l2.bias, l3.weight, l3.bias

    def forward(self, positions):
        shape = positions.shape
        if self.fused and positions.is_cuda:
            return ops.vertex_front(positions.contiguous().view(shape[0], shape[1], -1), None, *self._weights())
        x = ops.nerf_embed(positions.contiguous().view(shape[0] * shape[1], -1))
        return self.model(x).view(shape[0], shape[1], -1)


class Mask_Encoder(nn.Module):
    def __init__(self, input_size):
        super(Mask_Encoder, self).__init__()
        self.model = nn.Sequential(nn.Embedding(4, input_size))

    def forward(self, mask):
        shape = mask.shape
        return self.model(mask.contiguous().view(-1, 1).long()).view(shape[0], shape[1], -1)


def vertex_features(positional_encoder, mask_encoder, vertices, mask, img_features=None):
    """`positional_encoder(vertices) + mask_encoder(mask) [+ img_features]` (vision/model.py:229-236) in one launch.
    vertices (B,N,3), mask (B,N,1) float tokens 0..3, img_features (B,N,S) or None -> (B,N,S)."""
    if not (positional_encoder.fused and vertices.is_cuda):
        out = positional_encoder(vertices) + mask_encoder(mask)
        return out if img_features is None else out + img_features
    return ops.vertex_front(vertices, mask, *positional_encoder._weights(), emb=mask_encoder.model[0].weight,
                            add=img_features)
