"""surface_vision_transformers_b200 -- B200-native (sm_100a) hot path of the Surface Vision Transformer.

Public API mirrors the reference (SD3004/surface-vision-transformers):
    SiT                         <- models/sit.py::SiT
    masked_patch_pretraining    <- models/mpp.py::masked_patch_pretraining
plus the pieces the north star adds around them: FusedAdamW / FusedSGD (optim), regression_loss (fused criterion), DataParallel (ddp),
gather_patches / index tables (gather), DevicePrefetcher (loader: overlapped host -> device batch staging),
and the formats either side of the path: preprocess_meshes / PatchedNpyDataset (data),
load_weights_imagenet / load_ssl_checkpoint (interop), fit / fit_mpp (trainer), GraphedInference /
GraphedTrainStep (graphs: CUDA-graph capture of the step).
"""
from .sit import SiT, Transformer  # noqa: F401
from .mpp import masked_patch_pretraining, get_mask_from_prob, prob_mask_like  # noqa: F401
from .optim import FusedAdamW, FusedSGD  # noqa: F401
from .loss import regression_loss, RegressionLoss  # noqa: F401
from .ddp import DataParallel  # noqa: F401
from .gather import gather_patches, load_index_table  # noqa: F401
from .loader import DevicePrefetcher  # noqa: F401
from .data import preprocess_meshes, PatchedNpyDataset  # noqa: F401
from .interop import load_weights_imagenet, load_ssl_checkpoint  # noqa: F401
from .trainer import fit, evaluate, fit_mpp  # noqa: F401
from .graphs import GraphedInference, GraphedTrainStep  # noqa: F401
