"""Row-slab sharding: plan geometry, and the halo exchange on 2-3 ranks over gloo (CPU) — both the free functions and
ShardedUpscaler's own exchange code (halo="nccl": the code bench.py times when asked for the NCCL data plane).  The per-slab
compute in these CPU tests is the oracle standing in for the kernels (test infrastructure); the GPU twins are
tests/test_gpu_sharding.py (direct NVLink/IPC data plane, 1 GPU and N GPUs) and bench.py --gpus N (parity printed per run)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import fsr1_b200 as F
import oracle_lib as ol


@pytest.mark.parametrize("in_h,out_h,world", [(1080, 2160, 8), (2160, 4320, 8), (1440, 2160, 4), (1661, 2160, 3),
                                               (17, 31, 2), (9, 9, 4), (16, 40, 8)])
def test_plan_geometry(in_h, out_h, world):
    econ = ol.easu_con(64, in_h, 128, out_h)
    plan = F.SlabPlan(in_h, out_h, world, econ)
    covered = []
    for r in range(world):
        y0, y1 = plan.out_rows(r)
        covered += list(range(y0, y1))
        n0, n1 = plan.needed_in_rows(r)
        if y1 > y0:
            e0, e1 = plan.easu_rows(r)
            first, last = F.api.easu_input_rows(econ, in_h, e0, e1)
            assert (n0, n1) == (first, last + 1)      # the plan and the C ABI agree
        sends, recvs = plan.transfers(r)
        own = plan.owned_in_rows(r)
        got = set(range(max(own[0], n0), min(own[1], n1)))
        for peer, a, b in recvs:
            assert (r, a, b) in plan.transfers(peer)[0]  # every recv has the matching send
            got |= set(range(a, b))
        assert got == set(range(n0, n1))               # owned + received = needed, exactly
    assert covered == list(range(out_h))


def test_cfg5_halo_is_two_rows_each_side():
    plan = F.SlabPlan(2160, 4320, 8, ol.easu_con(3840, 2160, 7680, 4320))
    assert plan.needed_in_rows(3) == (808, 1082) and plan.owned_in_rows(3) == (810, 1080)
    assert plan.halo_bytes(3, 3840, 8) == 2 * 2 * 3840 * 8   # 2 x 61 440 B, SURVEY.md §8(e)
    assert plan.halo_bytes(0, 3840, 8) == 2 * 3840 * 8


def _worker(rank, world, port, shape, tmpdir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    iw, ih, ow, oh = shape
    frame = F.uniform(iw, ih, 4242)
    econ, rcon = ol.easu_con(iw, ih, ow, oh), ol.rcas_con(0.25)
    plan = F.SlabPlan(ih, oh, world, econ)
    own0, own1 = plan.owned_in_rows(rank)
    n0, n1 = plan.needed_in_rows(rank)
    owned = torch.from_numpy(frame[own0:own1].copy())
    window = torch.full((n1 - n0, iw, 4), float("nan"))
    F.exchange_halo(plan, rank, owned, window)
    assert np.array_equal(window.numpy(), frame[n0:n1])          # halo rows arrived where they belong
    # slab compute (oracle stand-in) on the window placed at its logical rows
    padded = np.zeros_like(frame)
    padded[n0:n1] = window.numpy()
    e0, e1 = plan.easu_rows(rank)
    y0, y1 = plan.out_rows(rank)
    tmp = ol.easu(padded, ow, oh, econ, y0=e0, y1=e1)
    out = ol.rcas(tmp, rcon, False, y0=y0, y1=y1)
    np.save(os.path.join(tmpdir, "slab%d.npy" % rank), out[y0:y1])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("shape,world", [((48, 40, 96, 80), 2), ((40, 33, 52, 43), 2), ((32, 30, 64, 60), 3)])
def test_halo_exchange_gloo(shape, world, tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, shape, str(tmp_path)), nprocs=world, join=True)
    iw, ih, ow, oh = shape
    frame = F.uniform(iw, ih, 4242)
    want = ol.rcas(ol.easu(frame, ow, oh), ol.rcas_con(0.25), False)
    got = np.concatenate([np.load(os.path.join(str(tmp_path), "slab%d.npy" % r)) for r in range(world)])
    assert np.array_equal(got, want)       # sharded == unsharded, bit for bit


def _worker_many(rank, world, port, shape, nframes):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    iw, ih, ow, oh = shape
    plan = F.SlabPlan(ih, oh, world, ol.easu_con(iw, ih, ow, oh))
    own0, own1 = plan.owned_in_rows(rank)
    n0, n1 = plan.needed_in_rows(rank)
    frames = [F.uniform(iw, ih, 100 + t) for t in range(nframes)]
    pairs = [(torch.from_numpy(f[own0:own1].copy()), torch.full((n1 - n0, iw, 4), float("nan"))) for f in frames]
    nops = F.exchange_halo_many(plan, rank, pairs)
    sends, recvs = plan.transfers(rank)
    assert nops == nframes * (len(sends) + len(recvs))
    for f, (_, window) in zip(frames, pairs):
        assert np.array_equal(window.numpy(), f[n0:n1])           # every frame's halo landed in ITS window
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("shape,world", [((48, 40, 96, 80), 2), ((32, 30, 64, 60), 3)])
def test_batched_halo_exchange_of_several_frames_gloo(shape, world):
    """exchange_halo_many: the halos of several frames in one batched group (what bench.py --halo-batch uses on NCCL)."""
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker_many, args=(world, port, shape, 3), nprocs=world, join=True)



def _worker_upscaler(rank, world, port, shape, tmpdir):
    """ShardedUpscaler(halo="nccl") on CPU tensors over gloo: its _exchange / exchange_many are what move the rows."""
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    iw, ih, ow, oh = shape
    nslots = 3
    frames = [F.uniform(iw, ih, 500 + t) for t in range(nslots)]
    up = F.ShardedUpscaler(iw, ih, ow, oh, world, rank, dtype=torch.float32, device="cpu", halo="nccl", slots=nslots)
    plan = up.plan
    o0, o1 = plan.owned_in_rows(rank)
    w0, w1 = plan.window_rows(rank)
    n0, n1 = plan.needed_in_rows(rank)
    for s in range(nslots):
        up.windows[s].fill_(float("nan"))
        up.input(s).copy_(torch.from_numpy(frames[s][o0:o1].copy()))
    nops = up._exchange(0)                       # one frame
    sends, recvs = plan.transfers(rank)
    assert nops == len(sends) + len(recvs)
    assert up.exchange_many([1, 2]) == 2 * nops  # two frames in one group
    for s in range(nslots):
        got = up.windows[s].numpy()[n0 - w0:n1 - w0]
        assert np.array_equal(got, frames[s][n0:n1]), "slot %d: halo rows did not land in the right window rows" % s
    # slab compute (oracle stand-in for the kernels) from the upscaler's own window, slot 2
    padded = np.zeros_like(frames[2])
    padded[w0:w1] = up.windows[2].numpy()
    e0, e1 = plan.easu_rows(rank)
    y0, y1 = plan.out_rows(rank)
    out = ol.rcas(ol.easu(padded, ow, oh, up.econ, y0=e0, y1=e1), up.rcon, False, y0=y0, y1=y1)
    np.save(os.path.join(tmpdir, "u%d.npy" % rank), out[y0:y1])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("shape,world", [((48, 40, 96, 80), 2), ((40, 33, 52, 43), 3)])
def test_sharded_upscaler_exchange_gloo(shape, world, tmp_path):
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_worker_upscaler, args=(world, port, shape, str(tmp_path)), nprocs=world, join=True)
    iw, ih, ow, oh = shape
    want = ol.rcas(ol.easu(F.uniform(iw, ih, 502), ow, oh), ol.rcas_con(0.25), False)
    got = np.concatenate([np.load(os.path.join(str(tmp_path), "u%d.npy" % r)) for r in range(world)])
    assert np.array_equal(got, want)


def test_plan_rejects_empty_slabs():
    with pytest.raises(ValueError):
        F.SlabPlan(5, 40, 8, ol.easu_con(64, 5, 128, 40))
