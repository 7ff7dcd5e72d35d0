"""Multi-GPU partitioning of the hot path: one process per GPU, streams are sticky to
``owner(stream) = stream_id mod world`` (a stream's frames must be processed in order on the GPU that holds
its temporal state), crops within a step are stateless.  The only exchange is the gather of the fixed-size
verdict records (72 B per stream) -- ``torch.distributed`` all_gather over NCCL/NVLink on GPUs, gloo on CPU
in the tests.  No other collective exists on this path and none is invented.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import _lib


class StreamSharder:
    def __init__(self, rank=None, world=None):
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world

    def owner(self, stream_id):
        return int(stream_id) % self.world

    def local_slot(self, stream_id):
        """Per-GPU state slot of a global stream id."""
        return int(stream_id) // self.world

    def select(self, stream_ids):
        """Indices (into this step's global batch) of the frames this rank owns, and their local slots."""
        sid = np.asarray(stream_ids, np.int64)
        idx = np.nonzero(sid % self.world == self.rank)[0]
        return idx, (sid[idx] // self.world).astype(np.int32)

    def shard_frames_round_robin(self, n_frames, crops_per_frame=None):
        """Stateless sharding for config 5 (frames with a variable number of boxes): greedy balance by crop
        count; returns the frame indices of this rank."""
        if crops_per_frame is None:
            return np.arange(self.rank, n_frames, self.world)
        order = np.argsort(-np.asarray(crops_per_frame), kind="stable")
        load = np.zeros(self.world, np.int64)
        mine = []
        for f in order:
            r = int(np.argmin(load))
            load[r] += max(int(crops_per_frame[f]), 1)
            if r == self.rank:
                mine.append(int(f))
        return np.array(sorted(mine), np.int64)

    def make_buffers(self, max_per_rank, device):
        """Preallocated gather buffers: ``send`` (max_per_rank records, padding slots pre-marked stream_id = -1) is meant to
        be passed as ``records_out`` to Engine.analyze_batch / capture_step so the vote kernel writes the records straight
        into the buffer the collective sends (no per-step allocation or copy); ``out`` receives world x max_per_rank records."""
        nbytes = max_per_rank * _lib.RECORD_BYTES
        send = torch.full((nbytes,), 0xFF, dtype=torch.uint8, device=device)
        out = torch.empty(self.world * nbytes, dtype=torch.uint8, device=device)
        return send, out

    def gather_records(self, records, max_per_rank, out=None, async_op=False):
        """records: uint8 tensor of n_local * 72 bytes (device for nccl, cpu for gloo).  Returns a
        (world * max_per_rank * 72) uint8 tensor on every rank; unused slots carry stream_id = -1.
        When ``records`` already is a full send buffer from ``make_buffers`` it is sent as is (no allocation, no copy)."""
        nbytes = max_per_rank * _lib.RECORD_BYTES
        if records.numel() == nbytes:
            send = records
        else:
            send = torch.full((nbytes,), 0xFF, dtype=torch.uint8, device=records.device)     # stream_id = -1 padding
            send[:records.numel()] = records
        if out is None:
            out = torch.empty(self.world * nbytes, dtype=torch.uint8, device=records.device)
        work = dist.all_gather_into_tensor(out, send, async_op=async_op)
        return (out, work) if async_op else out

    def globalize(self, gathered, max_per_rank):
        """Gathered records as a NumPy array with GLOBAL stream ids (an engine numbers its streams by local slot:
        global id = slot * world + owning rank, the inverse of ``select``)."""
        from .engine import RECORD_DTYPE
        rec = gathered.cpu().numpy().view(RECORD_DTYPE).copy().reshape(self.world, max_per_rank)
        for r in range(self.world):
            ok = rec[r]["stream_id"] >= 0
            rec[r]["stream_id"][ok] = rec[r]["stream_id"][ok] * self.world + r
        rec = rec.reshape(-1)
        return rec[rec["stream_id"] >= 0]

    @staticmethod
    def records_to_numpy(buf):
        from .engine import RECORD_DTYPE
        rec = buf.cpu().numpy().view(RECORD_DTYPE)
        return rec[rec["stream_id"] >= 0]
