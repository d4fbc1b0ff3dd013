"""Multi-rank host logic on CPU: world_size-2 gloo, no GPU.  Ranks own contiguous page ranges
(no data-path collective); only the per-block size index is gathered."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
torch = pytest.importorskip("torch")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_blocks, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    import oracle
    from csnappy_b200 import shard, synth

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, last = shard.block_range(n_blocks, rank, world)
    # each rank generates exactly its slice of the global batch (bench.py's weak-scaling layout)
    pages = synth.mixed_pages(last - first, 4096, seed=0x5EED0001, device="cpu", first_page=first, pool_bytes=1 << 20)
    _, sizes, _ = oracle.batch_compress(pages.numpy().reshape(-1, 4096), 13, "port")
    gathered = shard.gather_sizes(torch.from_numpy(sizes.astype(np.int32)), rank, world)
    if rank == 0:
        q.put((gathered.numpy().tolist(), shard.packed_offsets(gathered.numpy()).tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_reproduce_the_single_process_index():
    import torch.multiprocessing as mp

    sys.path.insert(0, ROOT)
    import oracle
    from csnappy_b200 import shard, synth

    n_blocks, world = 301, 2  # odd on purpose: ranks own 150 and 151 pages
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_blocks, q)) for r in range(world)]
    for p in procs:
        p.start()
    sizes, offsets = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pages = synth.mixed_pages(n_blocks, 4096, seed=0x5EED0001, device="cpu", pool_bytes=1 << 20)
    _, want, _ = oracle.batch_compress(pages.numpy().reshape(-1, 4096), 13, "port")
    # text and zero pages are position independent, random pages are seeded per rank slice: compare classes
    cls, _ = synth.page_classes(n_blocks, 0x5EED0001)
    det = cls != 2
    assert len(sizes) == n_blocks and (np.asarray(sizes)[det] == want[det]).all()
    assert (np.asarray(sizes)[~det] >= 4096).all()  # random pages do not shrink
    assert offsets[0] == 0 and offsets[-1] == sum(sizes) and len(offsets) == n_blocks + 1


def test_block_ranges_partition_the_batch():
    sys.path.insert(0, ROOT)
    from csnappy_b200 import shard

    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 4, 8):
            r = [shard.block_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    sizes = np.array([10, 10, 10, 1000, 10, 10, 1000, 10], dtype=np.int64)
    rr = shard.byte_balanced_ranges(sizes, 2)
    assert rr[0][0] == 0 and rr[-1][1] == len(sizes) and rr[0][1] == rr[1][0]
    with pytest.raises(ValueError):
        shard.block_range(10, 2, 2)
