"""Drop-in GCN_layer / GCN with the reference's constructor arguments, parameter names and shapes
(`weight (1,in,out)`, `bias (out,)` -- reference checkpoints load unchanged) and forward signatures:

    GCN_layer.forward(features, adj, activation)   pterotactyl/reconstruction/vision/model.py:335-363
    GCN.forward(features, adj_info)                 pterotactyl/reconstruction/vision/model.py:290-331

(the verbatim copies at reconstruction/autoencoder/model.py:96-124 and policies/DDQN/model.py:132-160
have the same interface).  `adj` stays the dense row-normalised tensor the callers already hold; its
CSR form is derived once and cached (graph.graph_of).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.parameter import Parameter

from . import ops


def _is_relu(fn):
    return fn is F.relu or fn is torch.relu or isinstance(fn, nn.ReLU)


class GCN_layer(nn.Module):
    def __init__(self, in_features, out_features, cut=0.33, do_cut=True):
        super(GCN_layer, self).__init__()
        self.weight = Parameter(torch.Tensor(1, in_features, out_features))
        self.bias = Parameter(torch.Tensor(out_features))
        self.reset_parameters()
        self.cut_size = cut
        self.do_cut = do_cut

    def reset_parameters(self):
        # same RNG consumption as the reference (model.py:345-349): weight first, then bias
        stdv = 6.0 / math.sqrt((self.weight.size(1) + self.weight.size(0)))
        stdv *= 0.3
        self.weight.data.uniform_(-stdv, stdv)
        self.bias.data.uniform_(-0.1, 0.1)

    def propagated(self):
        """Number of leading channels that go through the adjacency (model.py:355)."""
        out = self.weight.shape[2]
        return round(out * self.cut_size) if self.do_cut else out

    def forward(self, features, adj, activation):
        relu = _is_relu(activation)
        out = ops.gcn_layer(features, self.weight, self.bias, adj, self.propagated(), relu)
        return out if relu else activation(out)


class GCN(nn.Module):
    def __init__(self, input_features, args, ignore_touch_matrix=False):
        super(GCN, self).__init__()
        self.ignore_touch_matrix = ignore_touch_matrix
        self.num_layers = args.num_GCN_layers
        hidden_values = [input_features] + [args.hidden_GCN_size for _ in range(self.num_layers - 1)] + [3]
        self.layers = nn.ModuleList(
            GCN_layer(hidden_values[i], hidden_values[i + 1], args.cut, do_cut=i < self.num_layers - 1)
            for i in range(self.num_layers))
        self.check_nan = False  # the reference traps NaNs with one host sync per layer (model.py:326)

    def forward(self, features, adj_info):
        adj = adj_info["origional"] if self.ignore_touch_matrix else adj_info["adj"]
        n = self.num_layers
        out = ops.gcn_stack(
            features, adj,
            [l.weight for l in self.layers], [l.bias for l in self.layers],
            [l.propagated() for l in self.layers], [i < n - 1 for i in range(n)])
        if self.check_nan and torch.isnan(out).any():  # single deferred check instead of n syncs
            raise FloatingPointError("NaN in GCN output")
        return out


class Encoder(nn.Module):
    """The autoencoder's mesh encoder (pterotactyl/reconstruction/autoencoder/model.py:45-92): `num_GCN_layers`
    GCN layers of width `hidden_GCN_size` (the last one propagates every channel and has no activation), max over
    the vertices, four-layer MLP.  Same constructor arguments, sub-module names and initialisation order as the
    reference; the GCN layers run as one fused stack and the max through ops.vertex_max."""

    def __init__(self, input_features, args):
        super(Encoder, self).__init__()
        self.num_layers = args.num_GCN_layers
        hidden_values = [input_features] + [args.hidden_GCN_size for _ in range(self.num_layers)]
        self.layers = nn.ModuleList(
            GCN_layer(hidden_values[i], hidden_values[i + 1], args.cut, do_cut=i < self.num_layers - 1)
            for i in range(self.num_layers))
        hidden_values = [args.hidden_GCN_size, 500, 400, 300, args.encoding_size]
        n = len(hidden_values) - 1
        self.mlp = nn.Sequential(*[
            nn.Sequential(nn.Linear(hidden_values[i], hidden_values[i + 1]), nn.ReLU()) if i < n - 1
            else nn.Sequential(nn.Linear(hidden_values[i], hidden_values[i + 1])) for i in range(n)])

    def forward(self, features, adj_info):
        n = self.num_layers
        features = ops.gcn_stack(features, adj_info["adj"], [l.weight for l in self.layers],
                                 [l.bias for l in self.layers], [l.propagated() for l in self.layers],
                                 [i < n - 1 for i in range(n)])
        return self.mlp(ops.vertex_max(features)[0])


class Graph_Model(nn.Module):
    """The DDQN policy's graph value network (pterotactyl/policies/DDQN/model.py:65-129): action-mask MLP,
    positional + mask embeddings, `args.layers` GCN layers (300 -> hidden_dim ... -> num_actions), max over the
    vertices.  Reference constructor arguments, names, initialisation order and `forward(obs, next=False)`."""

    def __init__(self, args, adj):
        super().__init__()
        from .encoders import Mask_Encoder, Positional_Encoder
        self.adj = adj["adj"]  # the tensor adj_init registered with the CSR cache (the reference copies it)
        self.args = args
        self.num_layers = args.layers
        input_size = 100
        self.action_model = nn.Sequential(
            nn.Sequential(nn.Linear(50, 200), nn.ReLU()),
            nn.Sequential(nn.Linear(200, 100), nn.ReLU()),
            nn.Sequential(nn.Linear(100, input_size)))
        self.positional_embedding = Positional_Encoder(input_size)
        self.mask_embedding = Mask_Encoder(input_size)
        hidden_sizes = [input_size * 3] + [args.hidden_dim for _ in range(args.layers - 1)] + [args.num_actions]
        self.layers = nn.ModuleList(
            GCN_layer(hidden_sizes[i], hidden_sizes[i + 1], cut=args.cut, do_cut=(i != self.num_layers - 1))
            for i in range(args.layers))

    def forward(self, obs, next=False):
        sfx = "_n" if next else ""
        dev = self.layers[0].weight.device
        action_embedding = self.action_model(obs["mask" + sfx].float().to(dev))
        mesh = obs["mesh" + sfx][:, :, :3].float().to(dev)
        mask = obs["mesh" + sfx][:, :, 3:].float().to(dev)
        action_embedding = action_embedding.unsqueeze(1).repeat(1, mesh.shape[1], 1)
        vertex_features = torch.cat((action_embedding, self.positional_embedding(mesh), self.mask_embedding(mask)),
                                    dim=-1)
        n = self.num_layers
        x = ops.gcn_stack(vertex_features, self.adj, [l.weight for l in self.layers], [l.bias for l in self.layers],
                          [l.propagated() for l in self.layers], [i != n - 1 for i in range(n)])
        return ops.vertex_max(x)[0]
