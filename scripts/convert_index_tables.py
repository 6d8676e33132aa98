"""Converts the reference's integer gather tables (utils/triangle_indices_ico_6_sub_ico_{1,2}.csv; header row =
patch ids, one row per vertex slot) into compact uint16 .npy data files of shape (V, N).  Data only -- no code."""
import os
import sys

import numpy as np

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "surface_vision_transformers_b200", "data")
os.makedirs(out, exist_ok=True)
for sub in (1, 2):
    a = np.loadtxt(os.path.join(ref, "utils", f"triangle_indices_ico_6_sub_ico_{sub}.csv"), delimiter=",", skiprows=1,
                   dtype=np.int64)
    assert a.min() >= 0 and a.max() == 40961, (a.min(), a.max())
    np.save(os.path.join(out, f"triangle_indices_ico_6_sub_ico_{sub}.npy"), a.astype(np.uint16))
    print(sub, a.shape)
