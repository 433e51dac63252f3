"""Host-buffer front end of the forward path: the call a user with images in (pinned) host memory makes.

``HostFeed.run(batches, fn)`` walks a sequence of host batches ``(x1, x2, h_matrix)``; batch *i+1* is copied
host->device on a side stream into the second of two device buffer sets while batch *i* is computed on the
caller's stream, so the PCIe transfer (96 MB for 16 pairs at 512x512) hides behind the ~10 ms of kernels
instead of preceding them.  Every batch is still copied exactly once, inside the caller's timed region;
``fn(x1, x2, h)`` receives device tensors and is ordered after its copy by an event.  Plumbing only
(streams, events, pinned memory): the reference does the same job with ``.to(device)`` in its test loop
(ywz/mywork/test3real.py:190-200).

8-bit transport: images exist as 8-bit samples until ``transforms.ToTensor()`` turns them into fp32 on the host
(compressai/datasets/utils.py:101-102, test3real.py:323).  A host batch whose images are ``[B,H,W,3] uint8`` tensors is
copied as such -- a quarter of the PCIe bytes -- and converted by the first kernel on the device
(``functional.images_from_uint8``: u8 / 255, ToTensor's arithmetic) into the ``[B,3,H,W]`` fp32 tensors ``fn`` receives.
With 8 ranks feeding from one NUMA node's pinned memory that copy, not the kernels, set the end-to-end scaling (r01).
"""
import torch

from . import functional as F


class HostFeed:
    def __init__(self, device, like):
        """``like``: one host batch (tuple of tensors) giving the shapes/dtypes of the device buffers."""
        self.device = device
        self.copy_stream = torch.cuda.Stream(device=device)
        self.sets = [tuple(torch.empty(t.shape, dtype=t.dtype, device=device) for t in like) for _ in range(2)]
        # uint8 [B,H,W,C] images are converted on the device into these fp32 [B,C,H,W] tensors (one set per slot)
        self.f32 = [tuple(torch.empty((t.shape[0], t.shape[3], t.shape[1], t.shape[2]), dtype=torch.float32, device=device)
                          if self._is_u8_image(t) else None for t in like) for _ in range(2)]
        self.copied = [torch.cuda.Event() for _ in range(2)]
        self.consumed = [torch.cuda.Event() for _ in range(2)]
        self.bytes_per_batch = sum(t.numel() * t.element_size() for t in like)

    @staticmethod
    def _is_u8_image(t):
        return t.dtype == torch.uint8 and t.dim() == 4

    def _inputs(self, slot):
        """Device tensors handed to ``fn``: fp32 images (converted from the uint8 copies where those were shipped)."""
        return tuple(F.images_from_uint8(d, out=f) if f is not None else d for d, f in zip(self.sets[slot], self.f32[slot]))

    def _copy(self, slot, batch, first_use):
        main = torch.cuda.current_stream(self.device)
        cs = self.copy_stream
        if first_use:
            cs.wait_stream(main)                 # order after whatever produced/used the buffers before run()
        else:
            cs.wait_event(self.consumed[slot])   # the forward that read this buffer set two batches ago
        with torch.cuda.stream(cs):
            for d, s in zip(self.sets[slot], batch):
                d.copy_(s, non_blocking=True)
            self.copied[slot].record(cs)

    def run(self, batches, fn):
        """batches: indexable sequence of host batches; fn(*device_tensors) is called once per batch, in order."""
        n = len(batches)
        main = torch.cuda.current_stream(self.device)
        if n == 0:
            return
        self._copy(0, batches[0], True)
        for i in range(n):
            slot = i & 1
            if i + 1 < n:
                self._copy(slot ^ 1, batches[i + 1], i == 0)
            main.wait_event(self.copied[slot])
            fn(*self._inputs(slot))
            self.consumed[slot].record(main)
