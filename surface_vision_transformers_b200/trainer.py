"""Training loop for the scan-age / birth-age regression with the semantics of tools/train.py:271-363 (SURVEY 8(f)
rank 2), arranged so that the device never waits for the host:

* batches come from ``PatchedNpyDataset.batches`` through ``DevicePrefetcher`` (copy of batch i+1 under step i);
* the per-iteration ``loss.item()`` / ``.cpu()`` calls of the reference (train.py:293-296) are replaced by device-side
  accumulators -- sum of losses, sum of |target - prediction| -- that are read ONCE per epoch;
* under ``torchrun`` the model is wrapped in ``DataParallel`` (flat-gradient all-reduce overlapped with backward), every
  rank iterates a disjoint slice of the same permutation, and the epoch statistics are all-reduced.

What is kept from the reference: MSE (or L1) criterion on ``outputs.squeeze()``, train MAE per epoch, validation every
``val_epoch`` epochs in ``eval()`` + ``no_grad``, best-validation-MAE bookkeeping with an optional ``checkpoint.pth``
holding ``model.state_dict()`` (train.py:355-363), the same scalar names for an optional TensorBoard-like writer.

``fit_mpp`` is the same arrangement for the masked-patch pre-training loop of tools/pretrain.py:303-389.
"""
import os

import torch
import torch.distributed as dist

from .ddp import DataParallel
from .loader import DevicePrefetcher
from .loss import RegressionLoss

__all__ = ["fit", "evaluate", "fit_mpp"]


def _dist_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def _all_reduce_(t):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t)
    return t


def evaluate(model, dataset, batch_size, device, l1loss=False):
    """Validation pass of tools/train.py:311-337: returns (sum of batch losses, MAE, predictions, targets) with the
    statistics AND the predictions / targets of all ranks combined (host tensors, dataset order)."""
    rank, world = _dist_info()
    criterion = RegressionLoss(l1loss)      # train.py:245-248, one launch for loss and its gradient
    net = model.module if isinstance(model, DataParallel) else model
    was_training = net.training
    net.eval()
    stats = torch.zeros(3, dtype=torch.float64, device=device)   # sum of batch losses, sum |err|, count
    preds, targets = [], []
    with torch.no_grad():
        for x, y in DevicePrefetcher(dataset.batches(batch_size, rank=rank, world=world), device):
            out = net(x)
            stats[0] += criterion(out.squeeze(-1) if out.dim() > 1 else out, y).double()
            stats[1] += (out.reshape(-1) - y).abs().sum().double()
            stats[2] += y.numel()
            preds.append(out.reshape(-1).clone())
            targets.append(y.clone())
    _all_reduce_(stats)
    s = stats.cpu()
    net.train(was_training)
    p_all = torch.cat(preds).cpu() if preds else torch.zeros(0)
    t_all = torch.cat(targets).cpu() if targets else torch.zeros(0)
    if world > 1:
        # every rank evaluated the strided slice [rank::world] of the split: put the pieces back in dataset order, so
        # that what rank 0 saves as preds_test.pt covers the whole validation set (train.py:355-359), not its shard
        pieces = [None] * world
        dist.all_gather_object(pieces, (p_all, t_all))
        n = sum(int(pp.numel()) for pp, _ in pieces)
        p_full, t_full = torch.empty(n), torch.empty(n)
        for r, (pp, tt) in enumerate(pieces):
            p_full[r::world] = pp
            t_full[r::world] = tt
        p_all, t_all = p_full, t_full
    return float(s[0]), float(s[1] / s[2].clamp(min=1)), p_all, t_all


def fit(model, optimizer, train_set, val_set=None, *, epochs, batch_size, val_batch_size=None, val_epoch=1, device=None,
        l1loss=False, save_dir=None, save_ckpt=False, writer=None, seed=0, scheduler=None, log=None):
    """Returns a dict with the per-epoch history and the best validation MAE / epoch.

    ``model``: a ``SiT`` (wrapped in ``DataParallel`` here when torch.distributed is initialised with > 1 rank).
    ``train_set`` / ``val_set``: ``PatchedNpyDataset``-like objects (``batches(batch_size, shuffle, generator, rank,
    world)``).  ``writer``: optional object with ``add_scalar(tag, value, step)`` (TensorBoard SummaryWriter)."""
    rank, world = _dist_info()
    device = torch.device(device) if device is not None else next(model.parameters()).device
    if device.type != "cuda":
        raise RuntimeError("fit() drives the CUDA hot path (no CPU fallback)")
    net = model
    if world > 1 and not isinstance(model, DataParallel):
        net = DataParallel(model)
    core = net.module if isinstance(net, DataParallel) else net
    criterion = RegressionLoss(l1loss)      # train.py:245-248, one launch for loss and its gradient
    gen = torch.Generator()
    history = dict(train_loss=[], train_mae=[], val_loss=[], val_mae=[], lr=[])
    best_mae, best_epoch = float("inf"), None
    for epoch in range(epochs):
        core.train()
        gen.manual_seed(seed + epoch)                       # the same permutation on every rank
        stats = torch.zeros(4, dtype=torch.float64, device=device)   # sum loss, batches, sum |err|, samples
        batches = train_set.batches(batch_size, shuffle=True, generator=gen, rank=rank, world=world)
        for x, y in DevicePrefetcher(batches, device):
            optimizer.zero_grad(set_to_none=True)
            out = net(x)
            loss = criterion(out.squeeze(-1) if out.dim() > 1 else out, y)
            loss.backward()
            optimizer.step()
            with torch.no_grad():                           # device-side bookkeeping: no host sync inside the epoch
                stats[0] += loss.detach().double()
                stats[1] += 1
                stats[2] += (out.detach().reshape(-1) - y).abs().sum().double()
                stats[3] += y.numel()
        if scheduler is not None:
            scheduler.step()
        _all_reduce_(stats)
        s = stats.cpu()                                     # the one host read per epoch
        train_loss, train_mae = float(s[0] / s[1].clamp(min=1)), float(s[2] / s[3].clamp(min=1))
        history["train_loss"].append(train_loss)
        history["train_mae"].append(train_mae)
        history["lr"].append(optimizer.param_groups[0]["lr"])
        if writer is not None and rank == 0:
            writer.add_scalar("loss/train", train_loss, epoch + 1)
            writer.add_scalar("mae/train", train_mae, epoch + 1)
        if log is not None and rank == 0:
            log(f"| Epoch - {epoch + 1} | Loss - {train_loss:.4f} | MAE - {train_mae:.4f} | LR - {history['lr'][-1]}")
        if val_set is not None and (epoch + 1) % val_epoch == 0:
            val_loss, val_mae, preds, targets = evaluate(net, val_set, val_batch_size or batch_size, device, l1loss)
            history["val_loss"].append((epoch + 1, val_loss))
            history["val_mae"].append((epoch + 1, val_mae))
            if writer is not None and rank == 0:
                writer.add_scalar("loss/val", val_loss, epoch + 1)
                writer.add_scalar("mae/val", val_mae, epoch + 1)
            if log is not None and rank == 0:
                log(f"| Validation | Epoch - {epoch + 1} | Loss - {val_loss:.4f} | MAE - {val_mae:.4f} |")
            if val_mae < best_mae:
                best_mae, best_epoch = val_mae, epoch + 1
                if save_dir is not None and rank == 0:
                    os.makedirs(save_dir, exist_ok=True)
                    torch.save(dict(preds=preds, targets=targets), os.path.join(save_dir, "preds_test.pt"))
                    if save_ckpt:
                        torch.save(core.state_dict(), os.path.join(save_dir, "checkpoint.pth"))   # train.py:361-363
    return dict(history=history, best_mae=best_mae, best_epoch=best_epoch)


def _mpp_epoch_loss(ssl, dataset, batch_size, device, rank, world, train, optimizer=None, generator=None):
    """Mean over iterations of the MPP loss (pretrain.py: running_loss / (i + 1)), all ranks combined."""
    stats = torch.zeros(2, dtype=torch.float64, device=device)   # sum of batch losses, batches
    batches = dataset.batches(batch_size, shuffle=train, generator=generator, rank=rank, world=world)
    for x, _ in DevicePrefetcher(batches, device):
        if train:
            optimizer.zero_grad(set_to_none=True)
            loss, _ = ssl(x)
            loss.backward()
            optimizer.step()
        else:
            loss, _ = ssl(x)
        with torch.no_grad():
            stats[0] += loss.detach().double()
            stats[1] += 1
    _all_reduce_(stats)
    s = stats.cpu()                                         # the one host read per pass
    return float(s[0] / s[1].clamp(min=1))


def fit_mpp(ssl, optimizer, train_set, val_set=None, *, epochs, batch_size, val_batch_size=None, val_epoch=1, device=None,
            save_dir=None, writer=None, seed=0, scheduler=None, log=None):
    """Masked-patch pre-training with the semantics of tools/pretrain.py:303-389: ``mpp_loss, _ = ssl(inputs)`` per
    iteration, epoch loss = mean of the iteration losses, validation every ``val_epoch`` epochs under ``ssl.eval()`` +
    ``no_grad`` (the masking stays active, as in the reference), and on every improvement of the validation loss the two
    checkpoints the reference writes -- ``encoder-best.pt`` (the SiT's ``state_dict``) and ``encoder-decoder-best.pt``
    (the whole module's), each as ``{'epoch', 'model_state_dict', 'optimizer_state_dict', 'loss'}``, which is the file
    format ``load_ssl_checkpoint`` reads back for fine-tuning (train.py:213-223).

    ``ssl``: a B200 ``masked_patch_pretraining`` (wrapped in ``DataParallel`` here under torchrun).  Returns a dict with
    the history, the best validation loss and its epoch."""
    rank, world = _dist_info()
    device = torch.device(device) if device is not None else next(ssl.parameters()).device
    if device.type != "cuda":
        raise RuntimeError("fit_mpp() drives the CUDA hot path (no CPU fallback)")
    net = ssl
    if world > 1 and not isinstance(ssl, DataParallel):
        net = DataParallel(ssl)
    core = net.module if isinstance(net, DataParallel) else net
    gen = torch.Generator()
    history = dict(train_loss=[], val_loss=[], lr=[])
    best_val, best_epoch = float("inf"), None
    for epoch in range(epochs):
        core.train()
        gen.manual_seed(seed + epoch)
        train_loss = _mpp_epoch_loss(net, train_set, batch_size, device, rank, world, True, optimizer, gen)
        if scheduler is not None:
            scheduler.step()
        history["train_loss"].append(train_loss)
        history["lr"].append(optimizer.param_groups[0]["lr"])
        if writer is not None and rank == 0:
            writer.add_scalar("loss/train", train_loss, epoch + 1)
        if log is not None and rank == 0:
            log(f"| Epoch - {epoch + 1} | Loss - {train_loss:.4f} | LR - {history['lr'][-1]}")
        if val_set is not None and (epoch + 1) % val_epoch == 0:
            core.eval()
            with torch.no_grad():
                val_loss = _mpp_epoch_loss(net, val_set, val_batch_size or batch_size, device, rank, world, False)
            core.train()
            history["val_loss"].append((epoch + 1, val_loss))
            if writer is not None and rank == 0:
                writer.add_scalar("loss/val", val_loss, epoch + 1)
            if log is not None and rank == 0:
                log(f"| Validation | Epoch - {epoch + 1} | Loss - {val_loss} |")
            if val_loss < best_val:
                best_val, best_epoch = val_loss, epoch + 1
                if save_dir is not None and rank == 0:
                    os.makedirs(save_dir, exist_ok=True)
                    for name, module in (("encoder-best.pt", core.transformer), ("encoder-decoder-best.pt", core)):
                        torch.save({"epoch": epoch + 1, "model_state_dict": module.state_dict(),
                                    "optimizer_state_dict": optimizer.state_dict(), "loss": train_loss},
                                   os.path.join(save_dir, name))   # pretrain.py:378-389
    return dict(history=history, best_val_loss=best_val, best_epoch=best_epoch)
