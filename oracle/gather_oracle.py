"""TEST INFRASTRUCTURE (oracle) -- numpy restatement of the offline patch gather,
/root/reference/tools/preprocessing.py:72-84:

    normalised = (data - means) / stds                                   (:72)
    for j in range(num_patches):
        idx = indices_mesh_triangles[str(j)].to_numpy()                   (:82)   column j of the CSV
        out[i, :, j, :] = normalised[2*i][:, idx]                         (:83)   left hemisphere
        out[i + S, :, j, :] = normalised[2*i + 1][:, idx]                 (:84)   right hemisphere

The index tables are the reference's utils/triangle_indices_ico_6_sub_ico_{1,2}.csv (exact integers).
"""
import numpy as np


def gather_patches(mesh, table):
    """mesh: (S, C, 40962) float; table: (V, N) int -> (S, C, N, V), out[s,c,j,v] = mesh[s,c,table[v,j]]."""
    S, C, _ = mesh.shape
    V, N = table.shape
    out = np.zeros((S, C, N, V), dtype=mesh.dtype)
    for j in range(N):
        out[:, :, j, :] = mesh[:, :, table[:, j]]
    return out


def preprocessing_layout(hemis, table):
    """hemis: (2S, C, 40962) ordered L0,R0,L1,R1,... as built at preprocessing.py:62-67.
    Returns (2S, C, N, V) with left hemispheres in rows [0,S) and right in [S,2S) (:83-84)."""
    S2 = hemis.shape[0]
    S = S2 // 2
    g = gather_patches(hemis, table)
    out = np.zeros_like(g)
    out[:S] = g[0::2]
    out[S:] = g[1::2]
    return out


def zscore(data, means, stds):
    """preprocessing.py:72 ; means/stds shaped (1, C, 1)."""
    C = data.shape[1]
    return (data - means.reshape(1, C, 1)) / stds.reshape(1, C, 1)
