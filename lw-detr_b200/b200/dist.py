"""Multi-GPU plumbing of the inference path: images shard across ranks as independent
replicas, so the only collective is ONE broadcast of the weights from rank 0 at init.  Backend nccl on the
GPUs (NVLink/NVSwitch), gloo in the CPU tests."""
import torch
import torch.distributed as dist


def broadcast_module_weights(module, src=0):
    """Broadcast every parameter and floating-point buffer of `module` from rank `src` as ONE flat tensor
    (one collective launch; xlarge = 118 M parameters = 472 MB fp32)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return 0
    tensors = [p.data for p in module.parameters()] + [b.data for b in module.buffers() if b.is_floating_point()]
    if not tensors:
        return 0
    dev, dt = tensors[0].device, torch.float32
    flat = torch.cat([t.reshape(-1).to(dt) for t in tensors]) if dist.get_rank() == src else \
        torch.empty(sum(t.numel() for t in tensors), device=dev, dtype=dt)
    dist.broadcast(flat, src=src)
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].reshape(t.shape).to(t.dtype))
        off += n
    return flat.numel()


def shard_range(n_items, rank, world):
    """Contiguous image shard of rank `rank` (DistributedSampler(shuffle=False)-style partition, main.py:224)."""
    per = (n_items + world - 1) // world
    lo = min(rank * per, n_items)
    return lo, min(lo + per, n_items)
