"""Weight interop.

* ``load_weights_imagenet`` -- the timm ViT -> SiT key remap of the reference (utils/utils.py:11-35): LayerNorms, fused
  QKV weight (the timm QKV *bias* is dropped: vit-pytorch's ``to_qkv`` has none), attention projection, MLP, and the
  final ``norm`` into ``mlp_head.0``.  Patch embedding, cls token and position embedding are NOT transferred (different
  geometry), exactly as in the reference.
* ``load_ssl_checkpoint`` -- loads the encoder of an MPP pre-training checkpoint into a SiT.  tools/pretrain.py:378-389
  saves ``{'model_state_dict': model.state_dict(), ...}``; tools/train.py:216 hands that wrapper dict straight to
  ``load_state_dict(strict=False)``, which silently loads nothing.  This helper unwraps it (and accepts a bare state
  dict, or one with the ``transformer.`` prefix of ``masked_patch_pretraining``).
"""
import torch

__all__ = ["load_weights_imagenet", "load_ssl_checkpoint"]

_BLOCK_MAP = (
    ("0.norm.weight", "norm1.weight"), ("0.norm.bias", "norm1.bias"),
    ("1.norm.weight", "norm2.weight"), ("1.norm.bias", "norm2.bias"),
    ("0.fn.to_qkv.weight", "attn.qkv.weight"),
    ("0.fn.to_out.0.weight", "attn.proj.weight"), ("0.fn.to_out.0.bias", "attn.proj.bias"),
    ("1.fn.net.0.weight", "mlp.fc1.weight"), ("1.fn.net.0.bias", "mlp.fc1.bias"),
    ("1.fn.net.3.weight", "mlp.fc2.weight"), ("1.fn.net.3.bias", "mlp.fc2.bias"),
)


def load_weights_imagenet(state_dict, state_dict_imagenet, nb_layers):
    """Same signature and result as utils/utils.py::load_weights_imagenet: returns ``state_dict`` with the encoder
    entries replaced by the timm ViT tensors."""
    state_dict["mlp_head.0.weight"] = state_dict_imagenet["norm.weight"].data
    state_dict["mlp_head.0.bias"] = state_dict_imagenet["norm.bias"].data
    for i in range(nb_layers):
        for ours, theirs in _BLOCK_MAP:
            state_dict[f"transformer.layers.{i}.{ours}"] = state_dict_imagenet[f"blocks.{i}.{theirs}"].data
    return state_dict


def load_ssl_checkpoint(model, checkpoint, strict=True):
    """``checkpoint``: path or loaded object.  Returns the (missing, unexpected) key lists of load_state_dict."""
    if isinstance(checkpoint, (str, bytes)) or hasattr(checkpoint, "__fspath__"):
        checkpoint = torch.load(checkpoint, map_location="cpu")
    sd = checkpoint.get("model_state_dict", checkpoint) if isinstance(checkpoint, dict) else checkpoint
    if any(k.startswith("transformer.transformer.") or k.startswith("transformer.to_patch_embedding.") for k in sd):
        # state dict of the masked_patch_pretraining wrapper: keep the SiT entries only
        sd = {k[len("transformer."):]: v for k, v in sd.items() if k.startswith("transformer.")}
    own = model.state_dict()
    if not strict:
        sd = {k: v for k, v in sd.items() if k in own and tuple(v.shape) == tuple(own[k].shape)}
    res = model.load_state_dict(sd, strict=strict)
    return list(res.missing_keys), list(res.unexpected_keys)
