"""Generates tests/golden/consumers.npz from the REFERENCE's other two GCN consumers, imported unmodified and run
on the CPU in fp32 with the dense adjacency of the finger graph (N = 1949, hub rows included):

    autoencoder Encoder   pterotactyl/reconstruction/autoencoder/model.py:45-92   (GCN stack, last layer without
                          cut or activation, max over vertices, MLP)
    DDQN Graph_Model      pterotactyl/policies/DDQN/model.py:65-129               (action MLP, positional + mask
                          embeddings, GCN stack down to num_actions, max over vertices)

    python oracle/make_golden_consumers.py

Parameters are NOT stored: the test rebuilds the mirrors under the same torch seed (same torch build in the image;
the mirrors draw their parameters in the reference's order), which pins the initialisation order too.  Stored:
outputs, every GCN-layer gradient, a sample of the other gradients, and parameter checksums.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

ENC_ARGS = dict(num_GCN_layers=3, hidden_GCN_size=40, cut=0.33, encoding_size=20)
GM_ARGS = dict(layers=3, hidden_dim=30, num_actions=50, cut=0.33)
ENC_SEED, GM_SEED, B = 21, 22, 2


def inputs(n):
    """Deterministic inputs shared with the test (numpy generator, not torch's)."""
    rng = np.random.default_rng(77)
    feats = rng.random((B, n, 50), dtype=np.float32)
    g_latent = rng.standard_normal((B, ENC_ARGS["encoding_size"]), dtype=np.float32)
    mesh = np.concatenate([(rng.random((B, n, 3), dtype=np.float32) - 0.5) * 0.6,
                           rng.integers(0, 4, (B, n, 1)).astype(np.float32)], axis=-1)
    action_mask = rng.integers(0, 2, (B, 50)).astype(np.float32)
    g_value = rng.standard_normal((B, GM_ARGS["num_actions"]), dtype=np.float32)
    return feats, g_latent, mesh, action_mask, g_value


def dense_adj():
    from ptk_b200.graph import Graph
    adj = np.load(os.path.join(OUT, "adjacency.npz"))
    return Graph.from_csr(adj["p_adj_rowptr"], adj["p_adj_col"], "cpu").dense()


def record(out, tag, model, sample):
    for k, v in model.named_parameters():
        out[f"{tag}_sum.{k}"] = np.array([v.detach().double().sum().item()])
        if ".layers." in "." + k or k in sample:
            out[f"{tag}_grad.{k}"] = v.grad.numpy()
        else:
            out[f"{tag}_gradsum.{k}"] = np.array([v.grad.double().abs().sum().item()])


def main():
    from oracle import make_golden
    make_golden.install_stubs({})
    from pterotactyl.reconstruction.autoencoder import model as ae_model
    from pterotactyl.policies.DDQN import model as ddqn_model
    adj = dense_adj()
    n = adj.shape[0]
    feats, g_latent, mesh, action_mask, g_value = inputs(n)
    out = {}

    torch.manual_seed(ENC_SEED)
    enc = ae_model.Encoder(50, types.SimpleNamespace(**ENC_ARGS))
    x = torch.from_numpy(feats).requires_grad_(True)
    latent = enc(x, {"adj": adj})
    (latent * torch.from_numpy(g_latent)).sum().backward()
    out["enc_latent"] = latent.detach().numpy()
    out["enc_gx"] = x.grad.numpy()[:, ::8]
    record(out, "enc", enc, {"mlp.0.0.weight", "mlp.3.0.weight", "mlp.3.0.bias"})
    print("encoder", latent.shape, float(latent.detach().abs().mean()))

    torch.manual_seed(GM_SEED)
    gm = ddqn_model.Graph_Model(types.SimpleNamespace(**GM_ARGS), {"adj": adj})
    obs = {"mesh": torch.from_numpy(mesh), "mask": torch.from_numpy(action_mask)}
    value = gm(obs)
    (value * torch.from_numpy(g_value)).sum().backward()
    out["gm_value"] = value.detach().numpy()
    record(out, "gm", gm, {"action_model.0.0.weight", "positional_embedding.model.4.weight",
                           "mask_embedding.model.0.weight"})
    print("graph model", value.shape, float(value.detach().abs().mean()))
    np.savez_compressed(os.path.join(OUT, "consumers.npz"), **out)


if __name__ == "__main__":
    main()
