"""GPU, >= 2 devices: config 4 through the REAL sharder with real engines under NCCL -- streams are owned by
stream_id mod world (sharding.StreamSharder.select), every rank analyses its own streams, the vote kernel writes into the
preallocated send buffer, all_gather over NCCL, and every gathered record is compared with a single-GPU engine that
processed ALL streams (bit-identical: an image's result does not depend on the batch or the GPU it is in)."""
import os
import socket

import numpy as np
import pytest
import torch

import dfd_b200  # noqa: F401

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import dfd_b200  # noqa: F401
    from dfd_b200 import synth
    from dfd_b200.engine import Engine
    from dfd_b200.sharding import StreamSharder
    n, steps = 22, 6                                  # 22 streams over `world` ranks (ragged for world = 4, 8)
    sd = synth.make_state_dict()
    sh = StreamSharder()
    ids = np.arange(n)
    idx, slots = sh.select(ids)
    per = (n + world - 1) // world
    eng = Engine(device=rank, max_streams=per, max_batch=per, max_crop=512, detection_threshold=0.55)
    eng.load_state_dict(sd)
    ref = None
    if rank == 0:                                     # single-GPU engine over all streams
        ref = Engine(device=0, max_streams=n, max_batch=n, max_crop=512, detection_threshold=0.55)
        ref.load_state_dict(sd)
    send, out = sh.make_buffers(per, torch.device("cuda", rank))
    rng = np.random.RandomState(4)                    # same seed on every rank: identical global inputs
    bases = [synth.make_frame(synth.FAMILIES[s % 5], 240, 320, rng) for s in range(n)]
    for dtype in ("bf16", "fp32"):
        eng.reset(-1)
        if ref is not None:
            ref.reset(-1)
        for t in range(steps):
            fr = np.stack([np.clip(b.astype(np.int16) + rng.randint(-2, 3, b.shape[:2] + (1,)), 0, 255).astype(np.uint8) for b in bases])
            bx = synth.make_boxes(n, 240, 320, rng, lo=60, hi=200)
            full = int(t % 3 == 0)
            send.fill_(0xFF)
            m = len(idx)
            eng.analyze_batch(torch.from_numpy(fr[idx]).cuda(), slots, [full] * m, bx[idx], np.arange(m, dtype=np.int32), dtype=dtype,
                              records_out=send[:m * 72])
            gathered = sh.gather_records(send, per, out=out)
            torch.cuda.synchronize()
            allrec = sh.globalize(gathered, per)
            assert sorted(allrec["stream_id"].tolist()) == list(range(n))
            if ref is not None:
                want, _, _ = ref.analyze_batch(torch.from_numpy(fr).cuda(), ids, [full] * n, bx, np.arange(n, dtype=np.int32), dtype=dtype)
                want = ref.records_to_numpy(want)
                got = allrec[np.argsort(allrec["stream_id"])]
                for f in ("verdict", "fake_count", "real_count", "history_len", "frame_count", "last_vote", "vote_input",
                          "temporal_average", "stability_score", "face_probability", "forensic_probability"):
                    assert np.array_equal(got[f], want[f]), (dtype, t, f)
    dist.barrier()
    eng.close()
    if ref is not None:
        ref.close()
    dist.destroy_process_group()
    q.put((rank, "ok"))


def test_config4_sharded_streams_nccl_gather_matches_single_gpu():
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    assert sorted(q.get(timeout=5) for _ in range(world)) == [(r, "ok") for r in range(world)]
