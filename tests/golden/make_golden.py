"""Generates tests/golden/*.npz from the REFERENCE's own models/sit.py and models/mpp.py (imported unmodified from
/root/reference, with oracle.vit_shim standing in for the absent vit_pytorch package).

Weights and inputs come from numpy RandomState streams (stable across versions), so the fixtures only need to
store the seed, the config and the reference's outputs / gradients.  Run from the repo root in the build
container:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_loader  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    "sit_cls": dict(cfg=dict(dim=128, depth=2, heads=2, mlp_dim=256, num_patches=12, num_vertices=10, num_channels=4,
                             num_classes=1, dim_head=64, pool="cls"), batch=3, seed=11, kind="sit"),
    "sit_mean": dict(cfg=dict(dim=128, depth=2, heads=2, mlp_dim=128, num_patches=20, num_vertices=7, num_channels=4,
                              num_classes=1, dim_head=64, pool="mean"), batch=2, seed=12, kind="sit"),
    "mpp": dict(cfg=dict(dim=128, depth=2, heads=2, mlp_dim=256, num_patches=16, num_vertices=9, num_channels=4,
                         num_classes=1, dim_head=64, pool="cls"), batch=4, seed=13, kind="mpp",
                mpp=dict(mask_prob=0.5, replace_prob=0.8, swap_prob=0.3)),
}


def seeded_state(module, seed):
    """Deterministic weights: every tensor of the state_dict from one numpy RandomState stream."""
    rs = np.random.RandomState(seed)
    sd = {}
    for k, v in module.state_dict().items():
        a = rs.standard_normal(tuple(v.shape)).astype(np.float32)
        if k.endswith("norm.weight") or k == "mlp_head.0.weight":
            a = 1.0 + 0.1 * a
        elif v.dim() >= 2 and "embedding" not in k and "token" not in k:
            a = a * (1.0 / np.sqrt(v.shape[-1]))
        else:
            a = a * 0.1 if (k.endswith("bias")) else a
        sd[k] = torch.from_numpy(np.asarray(a, dtype=np.float32))
    return sd


def seeded_input(cfg, batch, seed):
    rs = np.random.RandomState(seed + 1000)
    x = rs.standard_normal((batch, cfg["num_channels"], cfg["num_patches"], cfg["num_vertices"])).astype(np.float32)
    y = (rs.rand(batch) * 19 + 26).astype(np.float32)
    return torch.from_numpy(x), torch.from_numpy(y)


def seeded_masks(cfg, batch, seed, mask_prob, replace_prob, swap_prob):
    """Fixed masks (numpy stream) with the reference's semantics: exactly ceil(p*N) masked tokens per sample."""
    import math
    rs = np.random.RandomState(seed + 2000)
    n = cfg["num_patches"]
    k = math.ceil(mask_prob * n)
    mask = np.zeros((batch, n), dtype=bool)
    for b in range(batch):
        mask[b, rs.permutation(n)[:k]] = True
    swap_sel = mask & (rs.rand(batch, n) < swap_prob / (1 - replace_prob))
    swap_src = rs.randint(0, n, size=(batch, n)).astype(np.int64)
    replace_sel = mask & (rs.rand(batch, n) < replace_prob)
    return mask, swap_sel, swap_src, replace_sel


def subsample(t):
    return t.detach().reshape(-1)[::7][:4096].numpy().copy()


def reference_mpp_forward_with_masks(ssl, batch, masks):
    """The reference's masked_patch_pretraining.forward (models/mpp.py:77-134) with its RNG calls replaced by the
    given masks -- executed through the reference module's own parameters and sub-modules."""
    import torch.nn.functional as F
    from einops import rearrange, repeat
    mask, swap_sel, swap_src, replace_sel = [torch.from_numpy(m) for m in masks]
    transformer = ssl.transformer
    batch = rearrange(batch, 'b c n v  -> b n (v c)')
    corrupted_batch = batch.clone().detach()
    randomized_input = corrupted_batch[torch.arange(corrupted_batch.shape[0]).unsqueeze(-1), swap_src]
    corrupted_batch[swap_sel] = randomized_input[swap_sel]
    corrupted_batch[replace_sel] = ssl.mask_token
    corrupted_batch = transformer.to_patch_embedding[-1](corrupted_batch)
    b, n, _ = corrupted_batch.shape
    cls_tokens = repeat(transformer.cls_token, '() n d -> b n d', b=b)
    corrupted_batch = torch.cat((cls_tokens, corrupted_batch), dim=1)
    corrupted_batch += transformer.pos_embedding[:, :(n + 1)]
    corrupted_batch = transformer.dropout(corrupted_batch)
    batch_out = transformer.transformer(corrupted_batch)
    batch_out = ssl.to_original(batch_out[:, 1:, :])
    return F.mse_loss(batch_out[mask], batch[mask]), batch_out


def main():
    SiT, MPP, _ = reference_loader.load_reference_models()
    for name, case in CASES.items():
        cfg, batch, seed = case["cfg"], case["batch"], case["seed"]
        torch.manual_seed(0)
        model = SiT(**cfg)
        out = {}
        x, y = seeded_input(cfg, batch, seed)
        if case["kind"] == "sit":
            model.load_state_dict(seeded_state(model, seed))
            pred = model(x)
            loss = torch.nn.functional.mse_loss(pred.squeeze(), y)
            loss.backward()
            out["pred"] = pred.detach().numpy()
            out["loss"] = np.float32(loss.item())
            named = dict(model.named_parameters())
        else:
            mp = case["mpp"]
            K = cfg["num_channels"] * cfg["num_vertices"]
            ssl = MPP(transformer=model, dim_in=cfg["dim"], dim_out=K, device="cpu", channels=cfg["num_channels"],
                      num_vertices=cfg["num_vertices"], **mp)
            ssl.load_state_dict(seeded_state(ssl, seed))
            masks = seeded_masks(cfg, batch, seed, **mp)
            # (a) the reference forward itself with its RNG, to pin loss semantics on a seeded run
            torch.manual_seed(seed)
            loss_rng, out_rng = ssl(x)
            out["loss_rng"] = np.float32(loss_rng.item())
            out["batch_out_rng_sub"] = subsample(out_rng)
            # (b) fixed masks, gradients
            loss, batch_out = reference_mpp_forward_with_masks(ssl, x, masks)
            loss.backward()
            out["loss"] = np.float32(loss.item())
            out["batch_out"] = batch_out.detach().numpy()
            for k, m in zip(("mask", "swap_sel", "swap_src", "replace_sel"), masks):
                out[k] = m
            named = dict(ssl.named_parameters())
        for k, p in named.items():
            out["grad_none/" + k] = np.array(p.grad is None)
            if p.grad is not None:
                out["grad_sub/" + k] = subsample(p.grad)
                out["grad_norm/" + k] = np.float32(p.grad.norm().item())
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in list(out.items())[:4]})


if __name__ == "__main__":
    main()


def make_gather_golden():
    """Patch gather exactly as tools/preprocessing.py:72-84 does it (pandas column str(j) of the reference CSV),
    on a numpy-seeded 2-subject (L,R,L,R) mesh; stores a SHA-256 of the float64 result plus a strided subsample."""
    import hashlib
    import pandas as pd
    out = {}
    for sub_ico, (num_patches, num_vertices) in {1: (80, 561), 2: (320, 153)}.items():
        df = pd.read_csv(os.path.join(reference_loader.REFERENCE_ROOT, "utils",
                                      f"triangle_indices_ico_6_sub_ico_{sub_ico}.csv"))
        rs = np.random.RandomState(100 + sub_ico)
        num_subjects, num_channels = 2, 4
        data = rs.standard_normal((num_subjects * 2, num_channels, 40962)).astype(np.float32)
        means = np.array([1.15, 0.037, 1.0, 0.07], dtype=np.float32).reshape(1, 4, 1)
        stds = np.array([0.41, 0.19, 0.39, 4.05], dtype=np.float32).reshape(1, 4, 1)
        normalised_data = (data - means.reshape(1, num_channels, 1)) / stds.reshape(1, num_channels, 1)   # :72
        res = np.zeros((num_subjects * 2, num_channels, num_patches, num_vertices))                        # :77
        for i in range(num_subjects):
            for j in range(num_patches):
                indices_to_extract = df[str(j)].to_numpy()                                                # :82
                res[i, :, j, :] = normalised_data[2 * i][:, indices_to_extract]                           # :83
                res[i + num_subjects, :, j, :] = normalised_data[2 * i + 1][:, indices_to_extract]        # :84
        out[f"sha256_f32/{sub_ico}"] = np.array(hashlib.sha256(res.astype(np.float32).tobytes()).hexdigest())
        out[f"sub/{sub_ico}"] = res.astype(np.float32).reshape(-1)[::997].copy()
        out[f"shape/{sub_ico}"] = np.array(res.shape)
    np.savez_compressed(os.path.join(HERE, "gather.npz"), **out)
    print("gather", {k: str(v)[:20] for k, v in out.items() if k.startswith("sha")})


if __name__ == "__main__":
    make_gather_golden()
