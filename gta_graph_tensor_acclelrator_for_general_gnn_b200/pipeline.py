"""Host-buffer front end: stream a sequence of feature batches through one layer program.

``execute()`` works on device tensors.  A caller whose features live in host memory pays a
PCIe copy in and out per batch (Reddit shape: 561 MB in, 119 MB out, about 2.4x the layer's
compute time), so the copies of batch i+1 / i-1 are overlapped with the kernels of batch i:
three CUDA streams (copy-in, compute, copy-out), ``depth`` device staging buffers, events for
the hand-offs.  Nothing here computes; it only orders copies and ``run`` calls.
"""
from __future__ import annotations

import torch

from . import kernels


def pinned_table(rows: int, width: int) -> torch.Tensor:
    """Pinned host ``[rows, width]`` fp32 table with the device row pitch (16-byte multiple), so
    the upload is one contiguous DMA instead of a pitched 2-D copy."""
    ld = kernels.pad4(width)
    buf = torch.zeros((max(rows, 1), ld), dtype=torch.float32).pin_memory()
    return buf[:rows, :width]


class HostPipeline:
    """``run(x_dev) -> y_dev`` applied to host batches with copy/compute overlap.

    submit(x_host, y_host) enqueues: H2D of x_host into a staging table, ``run`` on the compute
    stream, D2H of the result into y_host.  ``x_host`` / ``y_host`` should be pinned
    (``pinned_table``) for the copies to be asynchronous.  Call ``finish()`` before reading the
    last results.
    """

    def __init__(self, run, rows: int, in_width: int, device, depth: int = 2):
        self.run = run
        self.device = torch.device(device)
        self.depth = depth
        self.s_in = torch.cuda.Stream(self.device)
        self.s_run = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self.stage = [kernels.alloc_table(rows, in_width, self.device, zero=True) for _ in range(depth)]
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]
        self.ev_run = [torch.cuda.Event() for _ in range(depth)]      # staging buffer free again
        self.ev_first = torch.cuda.Event(enable_timing=True)
        self.ev_last = torch.cuda.Event(enable_timing=True)
        self.count = 0
        # results whose copy-out is still in flight: (tensor, event after its D2H).  They are kept alive HERE and
        # dropped only once that event has completed -- not handed to the caching allocator with record_stream(),
        # whose deferred frees made the steady state allocate now and then (cudaMalloc inside the timed region:
        # e2e steps of 20 ms instead of 11 in one run out of three)
        self.inflight = []
        cur = torch.cuda.current_stream(self.device)
        for s in (self.s_in, self.s_run, self.s_out):
            s.wait_stream(cur)

    def submit(self, x_host: torch.Tensor, y_host: torch.Tensor) -> None:
        slot = self.count % self.depth
        stage = self.stage[slot]
        with torch.cuda.stream(self.s_in):
            if self.count == 0:
                self.ev_first.record(self.s_in)
            if self.count >= self.depth:
                self.s_in.wait_event(self.ev_run[slot])          # the batch that used this slot has run
            same_pitch = x_host.stride(0) == stage.stride(0) and x_host.shape == stage.shape
            if same_pitch:      # one contiguous DMA over the padded rows
                flat_src = torch.as_strided(x_host, (x_host.shape[0] * x_host.stride(0),), (1,))
                flat_dst = torch.as_strided(stage, (stage.shape[0] * stage.stride(0),), (1,))
                flat_dst.copy_(flat_src, non_blocking=True)
            else:
                stage.copy_(x_host, non_blocking=True)
            self.ev_in[slot].record(self.s_in)
        with torch.cuda.stream(self.s_run):
            self.s_run.wait_event(self.ev_in[slot])
            y = self.run(stage)
            self.ev_run[slot].record(self.s_run)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_run[slot])
            y_host.copy_(y, non_blocking=True)
            self.ev_last.record(self.s_out)
            done = torch.cuda.Event()
            done.record(self.s_out)
        self.inflight.append((y, done))
        while len(self.inflight) > self.depth:
            _, ev = self.inflight.pop(0)
            ev.synchronize()          # two steps old: complete in the steady state
        self.count += 1

    def finish(self) -> float:
        """Wait for everything submitted; returns device milliseconds from the first copy-in to
        the last copy-out."""
        self.s_out.synchronize()
        self.s_run.synchronize()
        self.s_in.synchronize()
        self.inflight.clear()
        ms = self.ev_first.elapsed_time(self.ev_last) if self.count else 0.0
        self.count = 0
        return ms
