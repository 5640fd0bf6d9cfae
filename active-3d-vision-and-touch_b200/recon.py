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


class GraphedStep:
    """One training step (forward + loss + backward + optimizer) captured in a CUDA graph and replayed.

    The reference's step (vision/train.py:120-157) is ~600 kernel launches of 5-150 us at batch 16, so launch
    gaps and Python overhead are a measurable share.  Every ptk_b200 op is capture-safe (no host
    synchronisation, workspaces from torch's graph-private pool), so the whole step can be replayed from one
    graph launch.  `step_fn()` must read its inputs from tensors that stay at fixed addresses (copy new batches
    into them with `.copy_()`), return the loss tensor, and include `optimizer.zero_grad(set_to_none=True)`,
    `backward()` and `optimizer.step()`; the optimizer has to be built with `capturable=True` (pass the learning
    rate as a 0-d device tensor to change it between replays).

    World > 1: put `reducer.finish()` (ptk_b200.dist.GradReducer) between `backward()` and `optimizer.step()` --
    the bucketed NCCL all-reduces the gradient hooks launch are captured with the step and every rank replays its
    own graph (2 x B200, 16 objects per GPU: 18.4 ms a step against 20.8 ms eager, gradients equal to 2e-7;
    tools/graph_ddp_check.py).  Drop the GraphedStep (`del`) before `destroy_process_group()`: a live graph that
    holds NCCL kernels keeps the communicator from shutting down.
    """

    def __init__(self, step_fn, warmup=3):
        self.step_fn = step_fn
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # warm-up off the default stream: lazy initialisation, caches, autotuning
            for _ in range(warmup):
                step_fn()
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = step_fn()

    def __call__(self):
        self.graph.replay()
        return self.loss
