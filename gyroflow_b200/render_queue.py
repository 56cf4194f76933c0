"""Frame-sharded render queue: the multi-GPU shape of the path.

Frames are independent units (FrameTransform::at_timestamp depends only on immutable params + the timestamp,
src/core/stabilization/frame_transform.rs:165), so frame i goes to rank i % world and there is NO data-path collective.
The only communication is one broadcast, at job start, of the tables every rank needs (KernelParams template,
lens coefficients, per-frame matrix tables / quaternion tracks) from rank 0 — NCCL over NVLink on GPUs, gloo in the CPU
tests.  The reference has no counterpart: its render queue runs whole jobs in parallel (rendering/render_queue.rs:550-612)
but each job is a sequential decode -> warp -> encode loop on one device (rendering/mod.rs:451).
"""
import ctypes as C

import numpy as np

from . import abi


def shard_frames(n_frames, world, rank):
    """Frame indices owned by `rank`: round-robin, so every rank sees the same mix of timestamps."""
    return list(range(rank, n_frames, world))


def frame_owner(frame, world):
    return frame % world


def params_to_tensor(p, torch):
    """KernelParams (368 B) as a uint8 tensor (for the broadcast)."""
    raw = C.string_at(C.byref(p), C.sizeof(abi.KernelParams))
    return torch.frombuffer(bytearray(raw), dtype=torch.uint8).clone()


def params_from_tensor(t):
    p = abi.KernelParams()
    raw = bytes(t.cpu().numpy().tobytes())
    C.memmove(C.byref(p), raw, C.sizeof(abi.KernelParams))
    return p


def broadcast_tables(params, matrices, dist, torch, device, src=0):
    """One-time broadcast of the job's tables from `src`.  `params`/`matrices` are only read on the source rank;
    every rank passes the expected matrices shape.  Returns (KernelParams, matrices tensor on `device`)."""
    rank = dist.get_rank()
    pt = params_to_tensor(params, torch).to(device) if rank == src else torch.empty(C.sizeof(abi.KernelParams), dtype=torch.uint8, device=device)
    shape = torch.tensor(list(matrices.shape) if rank == src else [0, 0, 0], dtype=torch.int64, device=device)
    dist.broadcast(shape, src=src)
    mt = (torch.as_tensor(np.ascontiguousarray(matrices, dtype=np.float32)).to(device) if rank == src
          else torch.empty(tuple(int(v) for v in shape.tolist()), dtype=torch.float32, device=device))
    dist.broadcast(pt, src=src)
    dist.broadcast(mt, src=src)
    return params_from_tensor(pt), mt


def gather_results(local, dist, torch, device, dst=0):
    """Collect {frame: value} dicts (small per-frame results such as checksums) on `dst`, restoring frame order."""
    world = dist.get_world_size()
    keys = torch.tensor(sorted(local), dtype=torch.int64, device=device)
    vals = torch.tensor([local[k] for k in sorted(local)], dtype=torch.int64, device=device)
    n = torch.tensor([keys.numel()], dtype=torch.int64, device=device)
    ns = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(ns, n)
    m = int(max(int(v.item()) for v in ns))
    pad = lambda t: torch.cat([t, torch.full((m - t.numel(),), -1, dtype=torch.int64, device=device)])
    ks = [torch.empty(m, dtype=torch.int64, device=device) for _ in range(world)]
    vs = [torch.empty(m, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(ks, pad(keys)); dist.all_gather(vs, pad(vals))
    out = {}
    for k, v in zip(ks, vs):
        for a, b in zip(k.tolist(), v.tolist()):
            if a >= 0: out[a] = b
    return dict(sorted(out.items()))
