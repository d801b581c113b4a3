"""Host logic of the pair-sharded multi-GPU path: block partition and the result gather, run over
gloo with world_size 2 (CPU). The GPU variant runs two ranks on the box's GPU and must reproduce the
single-process result bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import LAUNCH_PARAMS, ROOT


def test_shard_range_partitions_every_count():
    from riv_slam_b200.sharding import shard_range, shard_segments
    for world in (1, 2, 4, 8):
        for n in (0, 1, 7, 8, 9, 4096, 1000):
            seen = []
            for r in range(world):
                b, e = shard_range(n, r, world)
                assert 0 <= b <= e <= n
                seen += list(range(b, e))
            assert seen == list(range(n))
    assert shard_segments(1001, 0, 8) == (0, 125) and shard_segments(1001, 7, 8) == (875, 1000)
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _fake_align(b, e):
    from riv_slam_b200.fast_apdgicp import RESULT_DTYPE
    out = np.zeros(e - b, dtype=RESULT_DTYPE)
    for i in range(b, e):
        out[i - b]["T"] = np.eye(4) * (i + 1)
        out[i - b]["fitness"] = 0.5 * i
        out[i - b]["iterations"] = i
        out[i - b]["converged"] = i % 2
    return out


def _gloo_worker(rank, world, port, n_pairs, use_gpu, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from riv_slam_b200.sharding import align_pairs_sharded
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        if use_gpu:
            from riv_slam_b200 import datagen
            from riv_slam_b200.fast_apdgicp import Handle, CloudSet, align_pairs
            pairs = [datagen.make_pair(4, 200 + i, n_src=800, n_tgt=900)[:2] for i in range(n_pairs)]
            H = Handle(0)
            H.set_params(**LAUNCH_PARAMS)

            def fn(b, e):
                if e == b:
                    from riv_slam_b200.fast_apdgicp import RESULT_DTYPE
                    return np.zeros(0, dtype=RESULT_DTYPE)
                return align_pairs(H, CloudSet(H, [p[0] for p in pairs[b:e]]), CloudSet(H, [p[1] for p in pairs[b:e]]))
            res = align_pairs_sharded(fn, n_pairs, rank, world)
        else:
            res = align_pairs_sharded(_fake_align, n_pairs, rank, world)
        q.put((rank, res.tobytes()))
    finally:
        dist.barrier()
        dist.destroy_process_group()


def _run(world, n_pairs, use_gpu, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, n_pairs, use_gpu, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return got


@pytest.mark.parametrize("n_pairs", [7, 8, 1])
def test_gather_over_gloo_world2(n_pairs):
    got = _run(2, n_pairs, False, 29611 + n_pairs)
    want = _fake_align(0, n_pairs).tobytes()
    assert got[0] == want and got[1] == want      # every rank ends with all records, in pair order


@pytest.mark.gpu
def test_sharded_equals_single_process_on_gpu():
    got2 = _run(2, 5, True, 29651)
    got1 = _run(1, 5, True, 29652)
    assert got2[0] == got2[1] == got1[0]
