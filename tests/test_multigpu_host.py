"""CPU tests of the N>1 host logic with world_size=2 over gloo: frame-parallel shards, row-band decomposition with
recomputed halos (checked with the oracle: assembled bands == full frame), and the max-over-ranks timing reduction
bench.py uses."""
import importlib.util
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import raisr_testlib as T

_spec = importlib.util.spec_from_file_location("raisr_sharding", os.path.join(T.PKG_DIR, "sharding.py"))
S = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(S)


def test_frame_shard_partition():
    for world in (1, 2, 4, 8):
        seen = sorted(f for r in range(world) for f in S.frame_shard(r, world, 37))
        assert seen == list(range(37))


def test_row_bands_cover_and_align():
    for out_h in (1080, 2160, 4320, 1082):
        for world in (1, 2, 4, 8):
            bands = S.row_bands(out_h, world)
            assert bands[0][0] == 0 and bands[-1][1] == out_h
            assert all(a[1] == b[0] for a, b in zip(bands, bands[1:]))
            assert all(r0 % 2 == 0 for r0, _ in bands)


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w, h, ratio = 160, 96, 2.0
        oW, oH = int(w * ratio), int(h * ratio)
        img = T.synth_frame(w, h, 8, seed=55)                      # every rank owns the same synthetic input
        m = T.OracleModel(T.filter_folder("filters_2x/filters_lowres"), 8)
        r0, r1 = S.row_bands(oH, world)[rank]
        a, b = S.band_input_rows(r0, r1, ratio, h, oH)
        # the rank only touches input rows [a, b): run the path on that slab and keep its own rows
        slab = T.oracle_process_y(img[a:b], oW, int((b - a) * ratio), m)
        mine = slab[r0 - int(a * ratio): r1 - int(a * ratio)]
        # frame-level borders come from the frame, not the slab: a band not touching the top/bottom edge has none
        t = torch.zeros((oH, oW), dtype=torch.uint8)
        t[r0:r1] = torch.from_numpy(mine.astype(np.uint8))
        dist.all_reduce(t, op=dist.ReduceOp.SUM)                   # bands are disjoint: SUM assembles the frame
        # max-over-ranks timing reduction as in bench.py
        el = torch.tensor([1.0 + rank], dtype=torch.float64)
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
        if rank == 0:
            full = T.oracle_process_y(img, oW, oH, m)
            np.save(os.path.join(tmp, "ok.npy"), np.array([int(np.array_equal(t.numpy(), full)), int(el.item() == world)]))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(not T.have_filters(), reason="trained filter folders not staged")
def test_row_band_shards_assemble_to_full_frame_gloo(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    ok = np.load(tmp_path / "ok.npy")
    assert ok[0] == 1, "assembled row bands differ from the full-frame result"
    assert ok[1] == 1, "max-over-ranks reduction"
