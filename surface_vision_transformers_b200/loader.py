"""Host -> device batch staging for the training loop (the part that feeds the hot path).

The reference keeps the whole dataset on the host and moves each mini-batch with ``inputs.to(device)`` inside the
loop (tools/train.py:281-283), so the 783 KB/sample H2D copy sits in front of every step.  ``DevicePrefetcher``
issues the copy of batch i+1 on a side stream while batch i is being computed; with pinned host tensors the copy
engine runs under the kernels and the step no longer waits for PCIe.

    for x, y in DevicePrefetcher(batches, device):      # batches: iterable of (pinned) host tensor tuples
        loss = criterion(model(x).squeeze(), y); loss.backward(); opt.step()

Device tensors handed out are views of an internal ring of ``depth`` staging buffers: a batch is valid until
``depth - 1`` further batches have been requested.
"""
import torch

__all__ = ["DevicePrefetcher"]


class DevicePrefetcher:
    def __init__(self, batches, device, depth=2):
        if depth < 2:
            raise ValueError("depth must be >= 2 (one buffer in use, one being filled)")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DevicePrefetcher stages batches for a CUDA device")
        self.batches = batches
        self.depth = depth
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.slots = [None] * depth        # per slot: tuple of device staging tensors
        self.filled = [None] * depth       # event: the copy into the slot has finished (recorded on the copy stream)
        self.released = [None] * depth     # event: the consumer is done with the slot (recorded on the compute stream)

    def _stage(self, slot, host_batch):
        if not isinstance(host_batch, (tuple, list)):
            host_batch = (host_batch,)
        bufs = self.slots[slot]
        if bufs is None or len(bufs) != len(host_batch) or any(
                b.shape != h.shape or b.dtype != h.dtype for b, h in zip(bufs, host_batch)):
            bufs = tuple(torch.empty(h.shape, dtype=h.dtype, device=self.device) for h in host_batch)
            self.slots[slot] = bufs
        with torch.cuda.stream(self.copy_stream):
            if self.released[slot] is not None:
                self.copy_stream.wait_event(self.released[slot])   # do not overwrite a batch that is still being read
            for b, h in zip(bufs, host_batch):
                b.copy_(h, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
            self.filled[slot] = ev

    def __iter__(self):
        it = iter(self.batches)
        compute = torch.cuda.current_stream(self.device)
        pending = []                      # slots staged but not yet handed out, in order
        nxt = 0
        try:
            for _ in range(self.depth - 1):
                self._stage(nxt, next(it))
                pending.append(nxt)
                nxt = (nxt + 1) % self.depth
        except StopIteration:
            pass
        prev = None
        while pending:
            slot = pending.pop(0)
            if prev is not None:
                # everything enqueued so far on the compute stream used the previous batch: mark it free
                ev = torch.cuda.Event()
                ev.record(compute)
                self.released[prev] = ev
            try:
                self._stage(nxt, next(it))
                pending.append(nxt)
                nxt = (nxt + 1) % self.depth
            except StopIteration:
                pass
            compute.wait_event(self.filled[slot])
            bufs = self.slots[slot]
            prev = slot
            yield bufs if len(bufs) > 1 else bufs[0]
