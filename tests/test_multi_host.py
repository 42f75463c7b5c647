"""The cross-rank exchange of pypercolate_b200.multi on CPU (gloo, world size 2):
word-wise integer all-reduce of the micro accumulators and the rank-ordered
Chan merge of canonical partials."""
import os
import socket

import numpy as np
import pytest

from pypercolate_b200 import multi


def test_shard_bounds_cover_everything():
    for total in (0, 1, 7, 100000):
        for world in (1, 2, 3, 8):
            spans = [multi.shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_chan_merge_equals_pooled_statistics():
    rng = np.random.RandomState(0)
    x = rng.rand(37, 5) * 1e6
    parts = np.split(x, [5, 6, 20])
    c, mu, s2 = 0, None, None
    for p in parts:
        c, mu, s2 = multi.chan_merge(c, mu, s2, len(p), p.mean(axis=0),
                                     ((p - p.mean(axis=0)) ** 2).sum(axis=0))
    assert c == 37
    np.testing.assert_allclose(mu, x.mean(axis=0), rtol=1e-13)
    np.testing.assert_allclose(s2, ((x - x.mean(axis=0)) ** 2).sum(axis=0), rtol=1e-11)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.RandomState(100 + rank)
    # limbs: 32-bit values in 64-bit words; include words near 2^32 * runs
    words = rng.randint(0, 2 ** 32, size=(50, 25)).astype(np.uint64) * np.uint64(1000 + rank)
    total = multi.allreduce_words(words)
    x = np.random.RandomState(7).rand(40, 3, 7)            # same on both ranks
    lo, hi = multi.shard_bounds(40, rank, world)
    mine = x[lo:hi]
    c, mu, s2 = multi.allgather_merge_canon(hi - lo, mine.mean(axis=0),
                                            ((mine - mine.mean(axis=0)) ** 2).sum(axis=0))
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), words=words, total=total, c=c, mu=mu, s2=s2)
    dist.destroy_process_group()


def test_gloo_world_size_2(tmp_path):
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(os.path.join(str(tmp_path), "r%d.npz" % i)) for i in range(world)]
    expect = (r[0]['words'].astype(object) + r[1]['words'].astype(object))
    assert np.array_equal(r[0]['total'].astype(object), expect)
    assert np.array_equal(r[0]['total'], r[1]['total'])
    x = np.random.RandomState(7).rand(40, 3, 7)
    for i in range(world):
        assert int(r[i]['c']) == 40
        np.testing.assert_allclose(r[i]['mu'], x.mean(axis=0), rtol=1e-13)
        np.testing.assert_allclose(r[i]['s2'], ((x - x.mean(axis=0)) ** 2).sum(axis=0), rtol=1e-11)
    # rank-ordered merge: bit-identical on every rank
    assert np.array_equal(r[0]['mu'], r[1]['mu']) and np.array_equal(r[0]['s2'], r[1]['s2'])


def test_communicator_id_travels_through_a_file(tmp_path):
    """The id of the C-ABI's NCCL communicator (pz_comm_unique_id) reaches the other ranks through
    a file on a shared file system -- no torch needed by a ctypes-only caller."""
    from pypercolate_b200 import _native
    path = str(tmp_path / "comm.id")
    with pytest.raises(RuntimeError):
        multi.exchange_comm_id_file(path, rank=1, timeout=0.05)      # nothing written yet
    cid = multi.exchange_comm_id_file(path, rank=0)
    assert len(cid) == _native.COMM_ID_BYTES
    assert multi.exchange_comm_id_file(path, rank=1) == cid
    assert multi.exchange_comm_id_file(path, rank=3) == cid
    assert _native.comm_unique_id() != cid                            # ids are fresh
