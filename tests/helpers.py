import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, GOLDEN)
from make_golden import CASES, seeded_input, seeded_masks, seeded_state, subsample  # noqa: E402,F401


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def rel_l2(a, b):
    a = torch.as_tensor(a).float().cpu()
    b = torch.as_tensor(b).float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def sit_kwargs(cfg):
    return dict(cfg)
