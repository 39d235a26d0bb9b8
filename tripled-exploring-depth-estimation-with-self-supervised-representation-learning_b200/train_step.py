"""Batch-sharded data-parallel training step around the fused loss ("DDP step hygiene").

The loss is per-image independent, so data parallelism is batch sharding with ONE exchange per step: the gradient
all-reduce.  The reference does that with mmcv's ``DistOptimizerHook`` -- after ``loss.backward()`` it flattens the
gradients into buckets, all-reduces them and divides by the world size (mono/core/utils/dist_utils.py:34-60), then
clips (``grad_clip=dict(max_norm=35)``, config/cfg_kitti_fm.py) and steps Adam (mono/apis/trainer.py:147-189).

``FlatGradStep`` is that step laid out for one B200 per rank:
  * every trainable parameter's ``.grad`` is a VIEW into one flat fp32 buffer, so the backward writes straight into it;
  * one ``all_reduce`` over the whole buffer per step (183 MB for cfg_kitti_fm = 0.5 ms on NVLink 5: there is nothing
    worth overlapping with a 30 ms step, hence no bucketing and no autograd hooks), optionally in bf16
    (``bf16_comm``: halves the bytes on the wire, fp32 accumulation in NCCL is not available so it is opt-in);
  * gradient clipping is fused on the flat buffer (one norm, one scale -- not one kernel pair per parameter), with the
    1/world averaging folded into the same scale;
  * fused multi-tensor Adam, capturable;
  * the WHOLE step -- forward, fused loss, backward, all-reduce, clip, Adam -- is captured into one CUDA graph
    (NCCL collectives are capturable), so the ~1500 small kernels of the eager step cost no host time at any rank
    count.  ``torch.nn.parallel.DistributedDataParallel`` is not used: its bucket hooks are what keep the eager
    multi-GPU step launch-bound.
SyncBatchNorm (cfg ``syncbn=True``) is NOT applied: its ~200 small collectives per step halve 2-GPU throughput
(measured in round 1); BatchNorm statistics are per rank (batch 8), as in single-GPU training.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist


class FlatGradStep:
    def __init__(self, model: torch.nn.Module, lr: float = 1e-4, max_norm: Optional[float] = 35.0, world: int = 1,
                 process_group=None, bf16_comm: bool = False, fused_adam: bool = True, capturable: bool = True):
        self.model, self.world, self.group, self.max_norm, self.bf16_comm = model, world, process_group, max_norm, bf16_comm
        self.params = [q for q in model.parameters() if q.requires_grad]
        dev = self.params[0].device
        total = sum(q.numel() for q in self.params)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        for q in self.params:                       # .grad of every parameter is a view of the flat buffer ...
            n = q.numel()
            dense = q.is_contiguous() or (q.dim() == 4 and q.is_contiguous(memory_format=torch.channels_last))
            if not dense:
                raise ValueError("FlatGradStep: parameters must be dense (contiguous or channels_last)")
            # ... with the parameter's own strides (channels_last conv weights), so that the fused multi-tensor Adam,
            # which walks parameter and gradient as flat memory, pairs the right elements
            q.grad = torch.as_strided(self.flat_grad, q.size(), q.stride(), off)
            off += n
        self.comm_buf = torch.empty(total, dtype=torch.bfloat16, device=dev) if (bf16_comm and world > 1) else None
        self.optim = torch.optim.Adam(self.params, lr=lr, weight_decay=0, capturable=capturable, fused=fused_adam or None)
        self.graph = None
        self.static_loss = None
        self.grad_norm = None

    def n_params(self):
        return self.flat_grad.numel()

    def broadcast_parameters(self, src=0):
        """Same initial weights on every rank (what DDP's constructor does)."""
        if self.world > 1:
            for t in list(self.model.parameters()) + list(self.model.buffers()):
                dist.broadcast(t.data, src, group=self.group)

    # ---------------------------------------------------------------------------------------------
    def reduce_and_clip(self):
        """all-reduce (sum) -> scale by min(1, max_norm / ||mean grad||) / world, on the flat buffer."""
        g = self.flat_grad
        if self.world > 1:
            if self.comm_buf is not None:
                self.comm_buf.copy_(g)
                dist.all_reduce(self.comm_buf, group=self.group)
                g.copy_(self.comm_buf)
            else:
                dist.all_reduce(g, group=self.group)
        inv_world = 1.0 / self.world
        if self.max_norm is not None:
            norm = torch.linalg.vector_norm(g) * inv_world                      # norm of the averaged gradient
            self.grad_norm = norm
            scale = torch.clamp(self.max_norm / (norm + 1e-6), max=1.0) * inv_world       # torch.nn.utils.clip_grad_norm_
            g.mul_(scale)
        elif self.world > 1:
            g.mul_(inv_world)

    def step(self, loss_fn: Callable[[], torch.Tensor]):
        """One optimisation step; ``loss_fn`` runs the forward and returns the scalar loss of this rank's shard."""
        self.flat_grad.zero_()
        loss = loss_fn()
        loss.backward()
        self.reduce_and_clip()
        self.optim.step()
        return loss

    # ---------------------------------------------------------------------------------------------
    def capture(self, loss_fn: Callable[[], torch.Tensor], warmup: int = 3):
        """Captures ``step(loss_fn)`` into a CUDA graph (static input tensors are the caller's responsibility)."""
        dev = self.flat_grad.device
        stream = torch.cuda.current_stream(dev)
        for _ in range(warmup):
            self.step(loss_fn)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=stream):
            self.static_loss = self.step(loss_fn)
        torch.cuda.synchronize(dev)
        return self

    def replay(self):
        self.graph.replay()
        return self.static_loss
