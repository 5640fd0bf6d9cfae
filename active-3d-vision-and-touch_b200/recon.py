"""The GCN-and-loss part of one reconstruction step, arranged as the reference arranges it.

`ChartDeformer` mirrors the three deformation iterations of Deformation.forward
(pterotactyl/reconstruction/vision/model.py:203-286): `mesh_deform_1` once, `mesh_deform_2` twice
with shared weights, only the vision-chart vertices are moved (model.py:250,270,283).  The vertex
feature encoders (NeRF positional MLP, mask embedding, CNN image pooling: model.py:27-164,367-414)
are the reference's own torch/cuDNN modules and out of this path's scope: they are passed in as a
callable `features(iteration, vertices) -> (B, N, input_size)`.

`recon_loss` is the loss of vision/train.py:141-144: loss_coeff * utils.chamfer_distance(...).mean().
"""
import torch
import torch.nn as nn

from . import utils
from .model import GCN


class ChartDeformer(nn.Module):
    def __init__(self, adj_info, args, input_size):
        super().__init__()
        self.adj_info = adj_info
        self.args = args
        self.mesh_deform_1 = GCN(input_size, args, ignore_touch_matrix=args.use_img)
        self.mesh_deform_2 = GCN(input_size, args)

    def forward(self, vision_charts, touch_charts, features):
        vc = vision_charts.shape[1]
        use_touch = self.args.use_touch and touch_charts is not None
        if use_touch and not self.args.use_img:
            vertices = torch.cat((vision_charts, touch_charts), dim=1)
        else:
            vertices = vision_charts
        update = self.mesh_deform_1(features(0, vertices), self.adj_info)
        vertices = torch.cat((vertices[:, :vc] + update[:, :vc], vertices[:, vc:]), dim=1)
        if use_touch and self.args.use_img:
            vertices = torch.cat((vertices, touch_charts), dim=1)
        for it in (1, 2):
            update = self.mesh_deform_2(features(it, vertices), self.adj_info)
            vertices = torch.cat((vertices[:, :vc] + update[:, :vc], vertices[:, vc:]), dim=1)
        return vertices


def recon_loss(vertices, faces, gt_points, number_points=10000, loss_coeff=9000.0, generator=None, uniforms=None):
    cd = utils.chamfer_distance(vertices, faces, gt_points, num=number_points, repeat=3, generator=generator,
                                uniforms=uniforms)
    return loss_coeff * cd.mean(), cd
