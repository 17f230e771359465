"""The multi-GPU data plane on hardware: fsr1_shard_* (direct NVLink / peer stores, sequence-number flow control).

  * several ranks in ONE process on ONE device (fsr1_shard_attach_local): the whole protocol — push, ready flags, EASU,
    credit flags, slot reuse, two-stream pipelining — runs on the single-GPU test box;
  * two PROCESSES attached through CUDA IPC handles (what bench.py --gpus N does): on one device when the box has one
    GPU, on two devices (NVLink peer access) when it has more;
  * the NCCL data plane (halo="nccl") on two devices when available.
Every result is compared bit-for-bit with the single-GPU api.upscale of the same frame, and that with the oracle."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import fsr1_b200 as F
import oracle_lib as ol

pytestmark = pytest.mark.gpu
api = F.api


def plain_upscale(frame_t, ow, oh, sharp=0.25):
    ih, iw = frame_t.shape[:2]
    # rows padded to a 16-byte multiple like the shard's own buffers, so both take the same (TMA-tiled) kernels
    tmp = torch.zeros((oh, (ow + 1) & ~1, 4), dtype=frame_t.dtype, device=frame_t.device)[:, :ow]
    out = torch.zeros((oh, (ow + 1) & ~1, 4), dtype=frame_t.dtype, device=frame_t.device)[:, :ow]
    api.upscale(frame_t, tmp, out, api.easu_con(iw, ih, iw, ih, ow, oh), api.rcas_con(sharp))
    return out


@pytest.mark.parametrize("shape,world", [((256, 144, 512, 288), 2), ((256, 150, 512, 300), 3), ((192, 144, 288, 216), 4),
                                         ((200, 120, 261, 157), 2)])
@pytest.mark.parametrize("one_stream", [False, True])
def test_shards_in_one_process_equal_single_gpu(shape, world, one_stream):
    iw, ih, ow, oh = shape
    nslots, nframes = 2, 7
    ups = [F.ShardedUpscaler(iw, ih, ow, oh, world, r, slots=nslots, halo="p2p", one_stream=one_stream, attach=False) for r in range(world)]
    for r, u in enumerate(ups):
        u.attach_local(ups[r - 1] if r > 0 else None, ups[r + 1] if r + 1 < world else None)
    frames = [torch.from_numpy(F.to_half(F.uniform(iw, ih, 900 + t))).cuda() for t in range(nframes)]
    want = [plain_upscale(f, ow, oh) for f in frames]
    torch.cuda.synchronize()
    s = torch.cuda.current_stream()
    got = []
    for i, fr in enumerate(frames):
        k = i % nslots
        if i >= nslots:                                  # collect the slot's previous frame before reusing it
            for u in ups:
                u.wait(k, s)
            got.append(torch.cat([u.output(k) for u in ups]).clone())
        for r, u in enumerate(ups):
            o0, o1 = u.plan.owned_in_rows(r)
            u.input(k).copy_(fr[o0:o1])
        for u in ups:                                     # rank order on the host; the device side is ordered by flags only
            u.submit(k, s)
    for i in range(nframes - nslots, nframes):
        for u in ups:
            u.wait(i % nslots, s)
        got.append(torch.cat([u.output(i % nslots) for u in ups]).clone())
    torch.cuda.synchronize()
    for u in ups:
        u.status()
    assert len(got) == nframes
    for i in range(nframes):
        assert torch.equal(got[i], want[i]), "frame %d differs from the single-GPU result" % i
    # and the single-GPU result is the reference's, within the fp16 tolerance
    src = frames[0].cpu().numpy().astype(np.float32)
    ref = ol.rcas(ol.easu(src, ow, oh), ol.rcas_con(0.25))
    assert np.abs(want[0].cpu().numpy().astype(np.float32) - ref)[..., :3].max() <= 1e-2
    for u in ups:
        u.close()


def test_shard_geometry_matches_python_plan_and_rejects_thin_slabs():
    for (ih, oh, world) in ((1080, 2160, 8), (2160, 4320, 8), (1440, 2160, 4), (1661, 2160, 3)):
        for r in (0, world // 2, world - 1):
            u = F.ShardedUpscaler(64, ih, 96, oh, world, r, slots=1, halo="p2p", attach=False)   # asserts SlabPlan == the C ABI's geometry inside
            assert u.info.halo_recv_bytes == u.plan.halo_bytes(r, 64, 8)
            u.close()
    with pytest.raises(F._lib.Fsr1Error):                # slabs of 1 input row cannot supply a 2-row halo: unsupported
        F.ShardedUpscaler(64, 8, 128, 16, 8, 3, halo="p2p", attach=False)


def _ipc_worker(rank, world, port, shape, ndev, tmpdir, halo):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch.distributed as dist
    dev = torch.device("cuda", rank % ndev)
    torch.cuda.set_device(dev)
    if halo == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)      # only carries the 64-byte IPC handles
    iw, ih, ow, oh = shape
    nslots, nframes = 2, 5
    up = F.ShardedUpscaler(iw, ih, ow, oh, world, rank, slots=nslots, halo=halo, device=dev)
    o0, o1 = up.plan.owned_in_rows(rank)
    s = torch.cuda.current_stream()
    outs = []
    for i in range(nframes):
        k = i % nslots
        if i >= nslots:
            up.wait(k, s)
            outs.append(up.output(k).clone())
        fr = torch.from_numpy(F.to_half(F.uniform(iw, ih, 700 + i))[o0:o1].copy()).to(dev)
        up.input(k).copy_(fr)
        up.submit(k, s)
    for i in range(nframes - nslots, nframes):
        up.wait(i % nslots, s)
        outs.append(up.output(i % nslots).clone())
    torch.cuda.synchronize()
    up.status()
    np.save(os.path.join(tmpdir, "slab%d.npy" % rank), torch.stack(outs).cpu().numpy().view(np.uint16))
    dist.barrier()
    up.close()
    dist.destroy_process_group()


def _run_ipc(shape, world, halo, tmp_path, ndev):
    port = 35500 + (os.getpid() % 2000)
    mp.spawn(_ipc_worker, args=(world, port, shape, ndev, str(tmp_path), halo), nprocs=world, join=True)
    iw, ih, ow, oh = shape
    got = np.concatenate([np.load(os.path.join(str(tmp_path), "slab%d.npy" % r)) for r in range(world)], axis=1)
    for i in range(got.shape[0]):
        want = plain_upscale(torch.from_numpy(F.to_half(F.uniform(iw, ih, 700 + i))).cuda(), ow, oh)
        assert np.array_equal(got[i], want.cpu().numpy().view(np.uint16)), "frame %d" % i


def test_two_processes_over_cuda_ipc(tmp_path):
    """Two ranks = two processes; the halo crosses the process boundary through a CUDA IPC mapping (and NVLink when the box
    has two GPUs; the same device otherwise)."""
    _run_ipc((320, 180, 640, 360), 2, "p2p", tmp_path, min(2, torch.cuda.device_count()))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpus_nccl_data_plane(tmp_path):
    _run_ipc((320, 180, 640, 360), 2, "nccl", tmp_path, 2)


@pytest.mark.skipif(torch.cuda.device_count() < 4, reason="needs 4 GPUs")
def test_four_gpus_p2p(tmp_path):
    _run_ipc((320, 200, 480, 300), 4, "p2p", tmp_path, 4)
