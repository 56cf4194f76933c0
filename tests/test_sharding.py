"""The N>1 path on CPU: world_size-2 gloo.  Tables are broadcast from rank 0, frames are sharded round-robin with no
data-path collective, per-frame results are gathered in frame order.  The warp itself is stood in for by the oracle
(tests may use it); what is under test is the sharding / broadcast / ordering logic that bench.py --gpus N runs over NCCL."""
import os
import socket
import zlib

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gyroflow_b200 import render_queue, synth
from tests import cases, np_producer, oracle_lib

N_FRAMES = 6


def _frame_crc(p, src, m):
    dst = np.zeros((p.output_height, p.output_stride), np.uint8)
    assert oracle_lib.undistort_image(src, dst, p, "RGBA8", "opencv_fisheye", None, m, None, 1) == 0
    return zlib.crc32(dst.tobytes())


def _tables():
    p = synth.base_kernel_params(96, 54)
    org, sm = cases.gyro()
    mats = np.stack([np_producer.frame_matrices(p, org, sm, 400.0 + 37.0 * i) for i in range(N_FRAMES)])
    p.matrix_count = mats.shape[1]
    return p, mats


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = torch.device("cpu")
    if rank == 0:
        p, mats = _tables()
    else:
        p, mats = synth.base_kernel_params(96, 54), np.zeros((0, 0, 0), np.float32)   # placeholders, overwritten by the broadcast
    p, mt = render_queue.broadcast_tables(p, mats, dist, torch, dev)
    assert p.matrix_count == 54 and tuple(mt.shape) == (N_FRAMES, 54, 14)
    src = synth.synthetic_frame(96, 54, "RGBA8", stride=p.stride)
    mine = {f: _frame_crc(p, src, mt[f].numpy()) for f in render_queue.shard_frames(N_FRAMES, world, rank)}
    allr = render_queue.gather_results(mine, dist, torch, dev)
    if rank == 0:
        q.put(allr)
    dist.barrier()
    dist.destroy_process_group()


def test_sharding_plan():
    for world in (1, 2, 3, 8):
        seen = sorted(f for r in range(world) for f in render_queue.shard_frames(17, world, r))
        assert seen == list(range(17))
        assert all(render_queue.frame_owner(f, world) == r for r in range(world) for f in render_queue.shard_frames(17, world, r))


def test_two_rank_gloo_matches_single_process():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs: pr.start()
    got = q.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    p, mats = _tables()
    src = synth.synthetic_frame(96, 54, "RGBA8", stride=p.stride)
    want = {f: _frame_crc(p, src, mats[f]) for f in range(N_FRAMES)}
    assert got == want
    assert len(set(want.values())) == N_FRAMES      # distinct timestamps really give distinct frames
