"""Patch gather: the on-device equivalent of tools/preprocessing.py:79-84.

``load_index_table(sub_ico)`` returns the reference's triangle index table
(utils/triangle_indices_ico_6_sub_ico_{1,2}.csv -- exact integers, shipped here as data/*.npy, converted by
scripts/convert_index_tables.py) as an int32 tensor of shape (V, N): column j lists the V ico-6 vertex ids of patch j.
"""
import os

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, vp

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
N_MESH_ICO6 = 40962


def load_index_table(sub_ico, device=None):
    path = os.path.join(_DATA, f"triangle_indices_ico_6_sub_ico_{sub_ico}.npy")
    if not os.path.exists(path):
        raise FileNotFoundError(f"no index table for sub_ico={sub_ico} (the reference ships sub_ico 1 and 2 only)")
    t = torch.from_numpy(np.load(path).astype(np.int32))
    return t.to(device) if device is not None else t


def gather_patches(mesh, table):
    """mesh (S,C,n_mesh) fp32 CUDA, table (V,N) int32 CUDA -> (S,C,N,V) fp32, out[s,c,j,v] = mesh[s,c,table[v,j]]
    (bit-exact copy; tools/preprocessing.py:83-84 without the L/R re-ordering, see preprocessing_layout)."""
    if not mesh.is_cuda:
        raise RuntimeError("gather_patches needs CUDA tensors (no CPU fallback)")
    mesh = mesh.contiguous().float()
    table = table.to(device=mesh.device, dtype=torch.int32).contiguous()
    S, C, n_mesh = mesh.shape
    V, N = table.shape
    out = torch.empty(S, C, N, V, dtype=torch.float32, device=mesh.device)
    lib = _lib.load()
    check(lib.svit_gather_patches(ptr(mesh), ptr(table), ptr(out), S, C, n_mesh, N, V,
                                  vp(torch.cuda.current_stream(mesh.device).cuda_stream)), "svit_gather_patches")
    return out


def preprocessing_layout(hemis_gathered):
    """(2S,...) ordered L0,R0,L1,R1,... -> left hemispheres first, then right (preprocessing.py:83-84)."""
    return torch.cat([hemis_gathered[0::2], hemis_gathered[1::2]], dim=0)
