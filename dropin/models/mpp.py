"""Drop-in for the reference's models/mpp.py: `from models.mpp import masked_patch_pretraining` (tools/pretrain.py)."""
from surface_vision_transformers_b200.mpp import (  # noqa: F401
    get_mask_from_prob,
    masked_patch_pretraining,
    prob_mask_like,
)
