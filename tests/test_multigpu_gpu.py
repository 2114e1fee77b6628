"""Multi-GPU row-band sharding on real devices (needs >= 2 GPUs; skipped otherwise): every rank computes its band of
the SAME frame through the C ABI on its own device, NCCL all_gather assembles the frame on every rank, and the result
must equal the single-GPU full-frame output bit for bit (BASELINE configs[3] scheme at reduced size)."""
import importlib.util
import os

import numpy as np
import pytest
import torch

import raisr_testlib as T

pytestmark = pytest.mark.gpu


def _load(name, file):
    spec = importlib.util.spec_from_file_location(name, os.path.join(T.PKG_DIR, file))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _worker(rank, world, port, tmp):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    B, S = _load("raisr_binding", "binding.py"), _load("raisr_sharding", "sharding.py")
    try:
        w, h, bits, ratio = 960, 544, 10, 2.0
        f = T.filter_folder("filters_2x/filters_denoise")
        img = T.synth_frame(w, h, bits, seed=404)
        oW, oH = int(w * ratio), int(h * ratio)
        eng = B.Engine(f, ratio, bits, T.VideoRange, 2, 2, device=rank)
        eng.set_res(w, h, oW, oH)
        d_in = torch.from_numpy(img.view(np.int16)).cuda()
        r0, r1 = S.row_bands(oH, world)[rank]
        band_rows = oH // world
        assert r1 - r0 == band_rows                      # equal bands so that all_gather applies
        d_out = torch.zeros((oH, oW), dtype=torch.int16, device="cuda")
        assert eng.process_device_rows(d_in.data_ptr(), d_in.stride(0) * 2, d_out.data_ptr(), d_out.stride(0) * 2, r0, r1) == 0
        torch.cuda.synchronize()
        parts = [torch.empty((band_rows, oW * 2), dtype=torch.uint8, device="cuda") for _ in range(world)]   # NCCL has no int16
        dist.all_gather(parts, d_out[r0:r1].contiguous().view(torch.uint8))
        frame = torch.cat(parts, 0).cpu().numpy().view(np.uint16)
        if rank == 0:
            full = np.zeros((oH, oW), np.uint16)
            e2 = B.Engine(f, ratio, bits, T.VideoRange, 2, 2, device=0)
            e2.set_res(w, h, oW, oH)
            assert e2.process_host(img, full) == 0
            e2.close()
            np.save(os.path.join(tmp, "ok.npy"), np.array([int(np.array_equal(frame, full))]))
        eng.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_row_band_shards_on_two_gpus(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, 29600 + os.getpid() % 1000, str(tmp_path)), nprocs=world, join=True)
    assert np.load(tmp_path / "ok.npy")[0] == 1


def _peer_worker(rank, world, port, tmp):
    """Peer-store row bands: every rank writes its band straight into rank 0's frame buffer (CUDA IPC handle, NVLink peer stores from
    the pass kernel); after a barrier rank 0 holds the whole frame -- no collective on the data path."""
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    B, S = _load("raisr_binding", "binding.py"), _load("raisr_sharding", "sharding.py")
    try:
        ok = True
        for (folder, bits, passes, mode) in (("filters_2x/filters_lowres", 8, 1, 1), ("filters_2x/filters_denoise", 10, 2, 2)):
            w, h, ratio = 960, 544, 2.0
            f = T.filter_folder(folder)
            img = T.synth_frame(w, h, bits, seed=505)
            oW, oH = int(w * ratio), int(h * ratio)
            bps = 1 if bits == 8 else 2
            tdt = torch.uint8 if bits == 8 else torch.int16
            eng = B.Engine(f, ratio, bits, T.VideoRange, passes, mode, device=rank)
            eng.set_res(w, h, oW, oH)
            d_in = torch.from_numpy(img.view(np.int16) if bits != 8 else img).cuda()
            junk = torch.empty(3 << 20, dtype=torch.uint8, device="cuda")          # make the frame buffer an INTERIOR pointer of a cached segment
            frame = torch.zeros((oH, oW), dtype=tdt, device="cuda") if rank == 0 else None
            handles = [B.ipc_export(frame.data_ptr()) if rank == 0 else None]
            dist.broadcast_object_list(handles, src=0)
            remote = frame.data_ptr() if rank == 0 else B.ipc_open(handles[0])
            r0, r1 = S.row_bands(oH, world)[rank]
            assert eng.process_device_rows(d_in.data_ptr(), d_in.stride(0) * bps, remote, oW * bps, r0, r1) == 0
            torch.cuda.synchronize()
            dist.barrier()
            if rank == 0:
                full = np.zeros((oH, oW), img.dtype)
                assert eng.process_host(img, full) == 0
                ok = ok and np.array_equal(frame.cpu().numpy().view(img.dtype), full)
            else:
                B.ipc_close(remote, handles[0])
            dist.barrier()
            eng.close()
            del junk
        if rank == 0:
            np.save(os.path.join(tmp, "ok.npy"), np.array([int(ok)]))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_row_bands_written_into_the_peer_frame_buffer(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_peer_worker, args=(world, 29700 + os.getpid() % 1000, str(tmp_path)), nprocs=world, join=True)
    assert np.load(tmp_path / "ok.npy")[0] == 1
