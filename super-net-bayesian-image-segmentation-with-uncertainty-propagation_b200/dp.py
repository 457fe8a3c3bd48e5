"""Batch-sharded data parallelism for the moment-propagation path (one process per GPU, torch.distributed).

The reference has no distributed code; this is the B200 scale-out of its two step functions:
  * inference / noise sweep / FGSM (Brats.py:742,999,1289,582-596): slices are independent -> shard the batch,
    replicate the weights, NO collective on the data path;
  * ELBO training (Brats.py:569-580): the NLL is a mean over B*HW (Brats.py:301-302,309) and the regularisers
    depend on weights only, so with shard sizes n_r the global gradient is sum_r (n_r / N) * grad(loss_r); one
    all-reduce(sum) of the flat fp32 gradient (7.76 M elements, 31 MB for BraTS) over NCCL/NVLink, bucketed in
    reverse layer order.  Keras' Adam(clipnorm=1.0) clips PER VARIABLE and is applied after the reduction
    (Brats.py:566,579).
Nothing here touches CUDA directly, so the logic is testable with the gloo backend on CPU.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

Tensor = torch.Tensor


def _parse_cpulist(text: str) -> List[int]:
    cpus: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def bind_to_gpu_numa_node(device_index: int) -> Optional[int]:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (sysfs), BEFORE any pinned host buffer is
    allocated, so the staging buffers of the host-to-host path are node-local.  With one process per GPU and no
    affinity every rank's pinned memory tends to land on one socket and the other socket's GPUs DMA across the
    inter-socket link.  Returns the node, or None when the topology is not visible (then nothing is changed)."""
    import os
    try:
        props = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def shard_bounds(global_batch: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[start, stop) of `rank`'s slices: per-GPU batch = ceil(global / world); the last ranks may
    get fewer (or zero) slices."""
    if global_batch < 0 or world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad shard arguments")
    per = -(-global_batch // world_size)
    start = min(rank * per, global_batch)
    return start, min(start + per, global_batch)


def shard_batch(x: Tensor, world_size: int, rank: int) -> Tensor:
    a, b = shard_bounds(x.shape[0], world_size, rank)
    return x[a:b]


def allreduce_gradients(params: Sequence[Tensor], local_count: int, global_count: int,
                        group: Optional[dist.ProcessGroup] = None, bucket_bytes: int = 8 << 20) -> None:
    """In-place: p.grad <- sum_r (n_r / N) p.grad_r, through flat buckets filled in REVERSE parameter order (the
    order backward produces them, so early buckets can overlap with the remaining wgrad kernels)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    if global_count <= 0:
        raise ValueError("global_count must be positive")
    scale = float(local_count) / float(global_count)
    todo = [p for p in reversed(list(params)) if p.grad is not None]
    bucket: List[Tensor] = []
    size = 0
    handles = []

    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([p.grad.reshape(-1) for p in bucket]).mul_(scale)
        h = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True)
        handles.append((h, flat, bucket))
        bucket, size = [], 0

    for p in todo:
        bucket.append(p)
        size += p.grad.numel() * p.grad.element_size()
        if size >= bucket_bytes:
            flush()
    flush()
    for h, flat, ps in handles:
        h.wait()
        off = 0
        for p in ps:
            n = p.grad.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
            off += n


def clip_by_norm_per_variable_(params: Iterable[Tensor], clipnorm: float = 1.0) -> None:
    """Keras optimizer `clipnorm`: every variable's gradient is rescaled to L2 norm <= clipnorm on its own
    (Brats.py:566), unlike torch's clip_grad_norm_ which uses the global norm.  Multi-tensor kernels: a handful of
    launches for all variables instead of three per variable."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    norms = torch.stack(torch._foreach_norm(grads))
    coef = torch.clamp(clipnorm / (norms + 1e-12), max=1.0)
    torch._foreach_mul_(grads, list(coef.unbind(0)))


def allreduce_flat_(flat: Tensor, local_count: int, global_count: int, group: Optional[dist.ProcessGroup] = None,
                    bucket_bytes: int = 8 << 20) -> None:
    """In-place weighted gradient all-reduce on ONE flat buffer (the FAST-mode engine's gradients are views of it):
    flat <- sum_r (n_r / N) flat_r, in bucket-sized slices so that NCCL pipelines them; no packing copies."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    if global_count <= 0:
        raise ValueError("global_count must be positive")
    flat.mul_(float(local_count) / float(global_count))
    step = max(1, bucket_bytes // flat.element_size())
    handles = [dist.all_reduce(flat[i:i + step], op=dist.ReduceOp.SUM, group=group, async_op=True)
               for i in range(0, flat.numel(), step)]
    for h in handles:
        h.wait()


def make_adam(params: Iterable[Tensor], lr: float = 1e-3) -> torch.optim.Adam:
    """tf.keras.optimizers.Adam defaults (beta 0.9/0.999, epsilon 1e-7) (Brats.py:566)."""
    params = list(params)
    fused = bool(params) and all(p.is_cuda for p in params)          # one multi-tensor kernel on the GPU
    return torch.optim.Adam(params, lr=lr, betas=(0.9, 0.999), eps=1e-7, fused=fused)


class DataParallelTrainer:
    """train_on_batch (Brats.py:569-580) over a sharded batch: forward/backward on the local slices, weighted
    gradient all-reduce, per-variable clipnorm, Adam."""

    def __init__(self, model, lr: float = 1e-3, kl_factor: float = 1e-5, clipnorm: float = 1.0,
                 group: Optional[dist.ProcessGroup] = None):
        self.model = model
        self.kl_factor = kl_factor
        self.clipnorm = clipnorm
        self.group = group
        self.opt = make_adam(model.parameters(), lr)

    def step(self, x_local: Tensor, y_local: Tensor, global_batch: int) -> Tensor:
        params = [p for p in self.model.parameters() if p.requires_grad]
        self.opt.zero_grad(set_to_none=True)
        n_local = x_local.shape[0]
        if getattr(self.model, "mode", "fp32") == "fast":
            # ONE collective sequence for every rank of a fast-mode job: fixed-size slices of a flat buffer in
            # parameter order.  A rank whose shard is empty (global_batch < world * ceil-split) contributes a zeroed
            # buffer of the same size, so the number and sizes of the all_reduce calls match across ranks.
            if n_local > 0:
                loss = self._fast_backward(x_local, y_local)
                flat = self._engine.flat_grad
            else:
                loss = torch.zeros((), device=params[0].device)
                flat = self._zero_flat_grad()
            allreduce_flat_(flat, n_local, global_batch, self.group)
        else:
            if n_local > 0:
                loss = self.model.elbo_loss(x_local, y_local, self.kl_factor)
                loss.backward()
            else:                       # an empty shard still takes part in the collective
                loss = torch.zeros((), device=params[0].device)
                for p in params:
                    p.grad = torch.zeros_like(p)
            allreduce_gradients(params, n_local, global_batch, self.group)
        clip_by_norm_per_variable_(params, self.clipnorm)
        self.opt.step()
        # engines re-derive their bf16 tensor-core operands when they see a new version (the training engine does it
        # inside its captured graph on every step)
        self.model._weights_version = getattr(self.model, "_weights_version", 0) + 1
        return loss.detach()

    _empty_flat = None

    def _zero_flat_grad(self) -> Tensor:
        """Zeroed flat gradient in the engine's parameter order (w_mu, w_sigma per layer of model.conv_names), with
        every p.grad a view of it: what an empty shard feeds the all-reduce."""
        m = self.model
        pairs = [getattr(m, name).weights() for name in m.conv_names]
        total = sum(w.numel() + ws.numel() for w, ws in pairs)
        if self._empty_flat is None or self._empty_flat.numel() != total:
            self._empty_flat = torch.zeros(total, device=pairs[0][0].device, dtype=torch.float32)
        flat = self._empty_flat.zero_()
        off = 0
        for w, ws in pairs:
            w.grad = flat[off:off + w.numel()].view_as(w)
            off += w.numel()
            ws.grad = flat[off:off + ws.numel()].view_as(ws)
            off += ws.numel()
        return flat

    _engine = None

    def _fast_backward(self, x: Tensor, y: Tensor) -> Tensor:
        """FAST mode: tensor-core forward, data-gradient and weight-gradient chain (engine.GradientEngine), then the
        regulariser terms of Brats.py:575-576, which depend on the weights only."""
        from .engine import GradientEngine
        m = self.model
        if self._engine is None or not self._engine.matches(x):
            self._engine = GradientEngine(m, x.shape[0], x.shape[1], x.shape[2], x.shape[3], x.device, train=True)
        loss, grads = self._engine.loss_and_weight_gradients(x, y, clip=(1e-12, 1e3), kl_factor=self.kl_factor)
        for name in m.conv_names:
            w, ws = getattr(m, name).weights()
            w.grad, ws.grad = grads[name]
        return loss[0]
