"""Host-side multi-rank logic (sharding, global canvas, detection gather) with world_size 2 over gloo on CPU."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from yolort_b200 import parallel


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 32, 33, 128):
        for world in (1, 2, 3, 8):
            b = [parallel.shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_global_canvas_is_the_unsharded_batch_shape():
    sizes = [(480, 640), (800, 600), (640, 427), (1080, 1920)]
    assert parallel.global_canvas(sizes, 640, 640) == (640, 640)
    assert parallel.global_canvas(sizes[3:], 640, 640) == (384, 640)   # a shard alone would pad differently
    assert parallel.global_canvas(sizes[3:], 640, 640, fixed_shape=(640, 640)) == (640, 640)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_images, D, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        boxes = torch.rand(n_images, D, 4, generator=g)
        scores = torch.rand(n_images, D, generator=g)
        labels = torch.randint(0, 80, (n_images, D), generator=g)
        counts = torch.randint(0, D + 1, (n_images,), generator=g, dtype=torch.int32)
        bounds = [parallel.shard_bounds(n_images, r, world) for r in range(world)]
        lo, hi = bounds[rank]
        packed = parallel.pack_detections(boxes[lo:hi], scores[lo:hi], labels[lo:hi])
        allp, allc = parallel.all_gather_detections(packed, counts[lo:hi], [b[1] - b[0] for b in bounds])
        dets = parallel.unpack_detections(allp, allc)
        ok = len(dets) == n_images
        for i, d in enumerate(dets):
            c = int(counts[i])
            ok &= torch.equal(d["boxes"], boxes[i, :c]) and torch.equal(d["scores"], scores[i, :c])
            ok &= torch.equal(d["labels"], labels[i, :c]) and d["labels"].dtype == torch.int64
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_gather_reassembles_the_full_ordered_list_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 5, 7, q)) for r in range(2)]   # uneven shards: 3 + 2
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
