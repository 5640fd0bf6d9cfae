"""The other two GCN consumers of the reference -- the autoencoder's Encoder and the DDQN Graph_Model -- through
their ptk_b200 mirrors (fused GCN stack, ops.vertex_max, fused positional embedding) against the reference's own
modules run on the CPU (tests/golden/consumers.npz from oracle/make_golden_consumers.py).  The mirrors are built under
the generator's torch seed: equal parameter checksums pin the reference's initialisation order as well."""
import types

import numpy as np
import pytest
import torch

import ptk_b200
from oracle import make_golden_consumers as mk

pytestmark = pytest.mark.gpu
TOL = 1e-5


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.fixture(scope="module")
def finger_graph(objects_dir):
    args = types.SimpleNamespace(use_touch=True, finger=True, num_grasps=5)
    adj_info, _ = ptk_b200.utils.load_mesh_vision(args, objects_dir + "/vision_charts.obj")
    return adj_info


def check_params_and_grads(g, tag, model):
    for k, v in model.named_parameters():
        if f"{tag}_grad.{k}" in g:
            assert rel_err(v.grad.cpu().numpy(), g[f"{tag}_grad.{k}"]) < TOL, k
        else:
            got = v.grad.double().abs().sum().item()
            assert abs(got - g[f"{tag}_gradsum.{k}"][0]) < 1e-4 * max(g[f"{tag}_gradsum.{k}"][0], 1e-30), k


def test_autoencoder_encoder_vs_reference_module(golden, finger_graph):
    g = golden("consumers")
    torch.manual_seed(mk.ENC_SEED)
    enc = ptk_b200.model.Encoder(50, types.SimpleNamespace(**mk.ENC_ARGS))
    for k, v in enc.named_parameters():  # same draws, same order as the reference's constructor
        assert v.detach().double().sum().item() == g[f"enc_sum.{k}"][0], k
    enc = enc.cuda()
    feats, g_latent, _, _, _ = mk.inputs(1949)
    x = torch.from_numpy(feats).cuda().requires_grad_(True)
    latent = enc(x, finger_graph)
    assert latent.shape == (mk.B, mk.ENC_ARGS["encoding_size"])
    assert rel_err(latent.detach().cpu().numpy(), g["enc_latent"]) < TOL
    (latent * torch.from_numpy(g_latent).cuda()).sum().backward()
    assert rel_err(x.grad.cpu().numpy()[:, ::8], g["enc_gx"]) < TOL
    check_params_and_grads(g, "enc", enc)


def test_ddqn_graph_model_vs_reference_module(golden, finger_graph):
    g = golden("consumers")
    torch.manual_seed(mk.GM_SEED)
    gm = ptk_b200.model.Graph_Model(types.SimpleNamespace(**mk.GM_ARGS), finger_graph)
    for k, v in gm.named_parameters():
        assert v.detach().double().sum().item() == g[f"gm_sum.{k}"][0], k
    gm = gm.cuda()
    _, _, mesh, action_mask, g_value = mk.inputs(1949)
    obs = {"mesh": torch.from_numpy(mesh), "mask": torch.from_numpy(action_mask),
           "mesh_n": torch.from_numpy(mesh), "mask_n": torch.from_numpy(action_mask)}
    value = gm(obs)
    assert value.shape == (mk.B, mk.GM_ARGS["num_actions"])
    assert rel_err(value.detach().cpu().numpy(), g["gm_value"]) < TOL
    assert torch.equal(gm(obs, next=True), value)
    (value * torch.from_numpy(g_value).cuda()).sum().backward()
    check_params_and_grads(g, "gm", gm)


def test_consumers_no_grad_forward_vs_reference_modules(golden, finger_graph):
    """Evaluation mode of both consumers (DDQN acts under torch.no_grad(), ddqn.py:81-95): the GCN stack then takes the
    tensor-core inference forward (ops.algo['fwd_infer']); same reference goldens, same 1e-5."""
    g = golden("consumers")
    torch.manual_seed(mk.ENC_SEED)
    enc = ptk_b200.model.Encoder(50, types.SimpleNamespace(**mk.ENC_ARGS)).cuda()
    feats, _, mesh, action_mask, _ = mk.inputs(1949)
    with torch.no_grad():
        latent = enc(torch.from_numpy(feats).cuda(), finger_graph)
    assert rel_err(latent.cpu().numpy(), g["enc_latent"]) < TOL
    torch.manual_seed(mk.GM_SEED)
    gm = ptk_b200.model.Graph_Model(types.SimpleNamespace(**mk.GM_ARGS), finger_graph).cuda()
    obs = {"mesh": torch.from_numpy(mesh), "mask": torch.from_numpy(action_mask)}
    with torch.no_grad():
        value = gm(obs)
    assert rel_err(value.cpu().numpy(), g["gm_value"]) < TOL
