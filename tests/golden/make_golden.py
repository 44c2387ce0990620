"""Generates tests/golden/grbm_checkpoints.npz from the reference's shipped checkpoints.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
The fixtures are the in-tree ground truth for the GRBM layout:
state-dict keys / dtypes / edge ordering / trained h, J of models/<QPU>_<k>_epochs/grbm.pth,
plus golden energies of three fixed spin patterns computed here in float64 directly from
static/eq6.png  (E = sum_i h_i s_i + sum_(ij) J_ij s_i s_j) -- no product or oracle code is
involved in producing them.  losses.json heads are kept as training-quality references.
"""
import json
import os
import sys

import numpy as np
import torch

REF = "/root/reference/models"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "grbm_checkpoints.npz")


def main():
    if not os.path.isdir(REF):
        sys.exit("reference checkpoints not found; fixtures are generated in the build container only")
    blob = {}
    meta = {}
    for name in sorted(os.listdir(REF)):
        sd = torch.load(os.path.join(REF, name, "grbm.pth"), map_location="cpu", weights_only=True)
        keys = {k: (str(v.dtype), list(v.shape)) for k, v in sd.items()}
        h = sd["_linear"].numpy()
        J = sd["_quadratic"].numpy()
        ei = sd["_edge_idx_i"].numpy()
        ej = sd["_edge_idx_j"].numpy()
        n = h.shape[0]
        idx = np.arange(n)
        pats = {
            "all_plus": np.ones(n),
            "even_plus": np.where(idx % 2 == 0, 1.0, -1.0),
            "mod3_plus": np.where(idx % 3 == 0, 1.0, -1.0),
        }
        h64, J64 = h.astype(np.float64), J.astype(np.float64)
        energies = {p: float(h64 @ s + (J64 * s[ei] * s[ej]).sum()) for p, s in pats.items()}
        dv = torch.load(os.path.join(REF, name, "dvae.pth"), map_location="cpu", weights_only=True)
        dvae_keys = {k: list(v.shape) for k, v in dv.items()}
        with open(os.path.join(REF, name, "parameters.json")) as f:
            params = json.load(f)
        with open(os.path.join(REF, name, "losses.json")) as f:
            losses = json.load(f)
        meta[name] = dict(keys=keys, dvae_keys=dvae_keys, energies=energies, sum_h=float(h64.sum()), sum_J=float(J64.sum()),
                          parameters=params, n_steps=len(losses["mse_losses"]),
                          mse_first=losses["mse_losses"][0], mse_last=losses["mse_losses"][-1],
                          dvae_first=losses["dvae_losses"][0])
        blob[name + "/linear"] = h
        blob[name + "/quadratic"] = J
        blob[name + "/edge_i"] = ei.astype(np.int32)
        blob[name + "/edge_j"] = ej.astype(np.int32)
    blob["meta_json"] = np.frombuffer(json.dumps(meta, sort_keys=True).encode(), dtype=np.uint8)
    np.savez_compressed(OUT, **blob)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
    for k, v in meta.items():
        print(k, v["energies"])


if __name__ == "__main__":
    main()
