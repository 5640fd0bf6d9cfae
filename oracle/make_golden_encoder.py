"""Generates tests/golden/encoder.npz from the REFERENCE's own Positional_Encoder / Mask_Encoder
(pterotactyl/reconstruction/vision/model.py:367-414), imported unmodified, run on the CPU in fp32.

    python oracle/make_golden_encoder.py

Pinned: the 63-wide NeRF embedding (model.py:381-391,396-397), the full positional MLP output, the mask embedding,
and the gradients of a fixed scalar loss with respect to the vertex positions and every parameter.  The encoder
parameters themselves are stored so the test loads identical weights (state-dict names are the reference's).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("PTK_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, REF)


def gsel_of(B, N, width):
    """Upstream gradient of the fixed scalar loss sum(feats * gsel); the test regenerates it from the same seed."""
    return np.random.default_rng(B * 1000003 + N * 101 + width).standard_normal((B, N, width), dtype=np.float32)


def main():
    from pterotactyl.reconstruction.vision import model as ref_model  # torch / numpy / PIL only

    torch.manual_seed(11)
    out = {}
    for tag, (B, N, width) in {"small": (2, 57, 48), "c3": (2, 487, 448)}.items():
        enc = ref_model.Positional_Encoder(width)
        menc = ref_model.Mask_Encoder(width)
        pos = ((torch.rand(B, N, 3) - 0.5) * 0.6).requires_grad_(True)  # chart vertices live in [-0.3, 0.3]^3
        mask = torch.randint(0, 4, (B, N, 1)).float()
        emb = enc.nerf_embedding(pos.reshape(B * N, 3))
        feats = enc(pos) + menc(mask)
        gsel = torch.from_numpy(gsel_of(B, N, width))
        (feats * gsel).sum().backward()
        out[f"{tag}_pos"] = pos.detach().numpy()
        out[f"{tag}_mask"] = mask.numpy()
        out[f"{tag}_embedding"] = emb.detach().numpy()          # (B*N, 60): sin/cos part only
        out[f"{tag}_feats"] = feats.detach().numpy()
        out[f"{tag}_gpos"] = pos.grad.numpy()
        for k, v in enc.state_dict().items():
            out[f"{tag}_enc.{k}"] = v.numpy()
        for k, v in enc.named_parameters():
            out[f"{tag}_gradenc.{k}"] = v.grad.numpy()
        for k, v in menc.state_dict().items():
            out[f"{tag}_menc.{k}"] = v.numpy()
        for k, v in menc.named_parameters():
            out[f"{tag}_gradmenc.{k}"] = v.grad.numpy()
        print(tag, emb.shape, feats.shape)
    np.savez_compressed(os.path.join(OUT, "encoder.npz"), **out)


if __name__ == "__main__":
    main()
