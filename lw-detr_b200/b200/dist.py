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
    with torch.no_grad():
        for t in list(module.parameters()) + [b for b in module.buffers() if b.is_floating_point()]:
            n = t.numel()
            t.copy_(flat[off:off + n].reshape(t.shape).to(t.dtype))      # on the tensor itself: bumps its _version, so an
            off += n                                                      # already packed engine is re-packed (LWDETR._weights_signature)
    if hasattr(module, "_engine_sig"):
        module._engine_sig = None
    return flat.numel()


# ------------------------------------------------------------------------------------------------ NCCL communicator for the C ABI
_NCCL = None


def _nccl_lib():
    """The libnccl torch already loaded (soname libnccl.so.2), through ctypes - no torch-private API."""
    global _NCCL
    if _NCCL is None:
        import ctypes
        import glob
        import os
        try:
            _NCCL = ctypes.CDLL("libnccl.so.2", mode=ctypes.RTLD_GLOBAL)
        except OSError:
            import nvidia.nccl
            cand = glob.glob(os.path.join(list(nvidia.nccl.__path__)[0], "lib", "libnccl.so*"))
            _NCCL = ctypes.CDLL(cand[0], mode=ctypes.RTLD_GLOBAL)
    return _NCCL


def nccl_comm_for_process_group(device):
    """A raw ncclComm_t (as int) spanning the ranks of the default process group, one rank per GPU: the unique id is made
    on rank 0 and shipped through the existing process group, the communicator by ncclCommInitRank.  For
    lwdetr_broadcast_weights (include/lwdetr_b200.h); keep the returned handle alive and destroy it with nccl_comm_destroy."""
    import ctypes
    nccl = _nccl_lib()
    rank, world = dist.get_rank(), dist.get_world_size()
    uid = (ctypes.c_byte * 128)()
    if rank == 0:
        rc = nccl.ncclGetUniqueId(ctypes.byref(uid))
        if rc != 0:
            raise RuntimeError("ncclGetUniqueId failed: %d" % rc)
    t = torch.tensor(list(bytes(uid)), dtype=torch.uint8, device=device)
    dist.broadcast(t, src=0)

    class _Uid(ctypes.Structure):
        _fields_ = [("internal", ctypes.c_byte * 128)]
    u = _Uid()
    ctypes.memmove(ctypes.byref(u), bytes(t.cpu().tolist()), 128)
    comm = ctypes.c_void_p()
    nccl.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, _Uid, ctypes.c_int]
    with torch.cuda.device(device):
        rc = nccl.ncclCommInitRank(ctypes.byref(comm), world, u, rank)
    if rc != 0:
        raise RuntimeError("ncclCommInitRank failed: %d" % rc)
    return comm.value


def nccl_comm_destroy(comm):
    import ctypes
    _nccl_lib().ncclCommDestroy(ctypes.c_void_p(comm))


def broadcast_engine_weights(engine, device, src=0):
    """SURVEY.md 8e as specified: ONE ncclBroadcast of the PACKED arena (16-bit weights + fp32 vectors; tiny 24 MB ... xlarge
    236 MB) through the C ABI.  Every rank must have packed weights of the same config first (values irrelevant off-root)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return 0
    comm = nccl_comm_for_process_group(device)
    try:
        engine.broadcast_weights(comm, root=src)
        torch.cuda.synchronize(device)
    finally:
        nccl_comm_destroy(comm)
    return engine.arena_bytes()


def shard_range(n_items, rank, world):
    """Contiguous image shard of rank `rank` (DistributedSampler(shuffle=False)-style partition, main.py:224)."""
    per = (n_items + world - 1) // world
    lo = min(rank * per, n_items)
    return lo, min(lo + per, n_items)
