#!/usr/bin/env python
"""Headline benchmark: frames/s end to end at 720p (BASELINE.json `metric`).

    python bench.py --gpus N --steps K --warmup W            # our arm (N>1: launched by torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

A step = one pass of the whole per-frame hot path (six forensic signals, face-crop preparation,
EfficientNet-B0, sigmoid + heuristics, vote-input selection, 10-frame vote) over one 1280x720
frame from each of S=256 streams on every GPU (weak scaling), one synthetic face box per frame,
plus -- for N>1 -- the NCCL gather of the per-stream verdict records.  Prints ONE JSON line.

Default classifier precision: fp32 accuracy mode on the tensor cores (3xTF32), the mode that meets the
reference-parity gate (|dp| <= 1e-4; measured 7e-6).  bf16 is reported beside it (`other_configs`).
`value`: frames resident in HBM.  `e2e`: the /analyze wire format -- JPEG streams in pinned host memory ->
H2D -> device JPEG decode -> path -> D2H of the verdict records (`e2e_raw`: the same with raw frames).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import dfd_b200  # noqa: E402
from dfd_b200 import synth  # noqa: E402

H, W = 720, 1280
STREAMS = 256          # frames per step per GPU
METRIC = "frames_per_sec_end_to_end_720p"
UNIT = "frames/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
def make_inputs(n_streams, n_sets, seed, jpeg=False):
    """n_sets x n_streams synthetic 720p BGR frames in pinned host memory + one face box per frame.
    jpeg=True: the frames are what cv2.imdecode returns for their JPEG streams (quality 85, 4:2:0 -- the /analyze wire
    format), and the streams are returned as well, so that `value` (resident frames) and `e2e` (JPEG ingest) run the
    path on identical pixels."""
    rng = np.random.RandomState(seed)
    bases = []
    for fam in synth.FAMILIES:
        for _ in range(2):
            bases.append(synth.make_frame(fam, H, W, rng))
    frames = torch.empty((n_sets, n_streams, H, W, 3), dtype=torch.uint8)
    if torch.cuda.is_available():
        frames = frames.pin_memory()
    fn = frames.numpy()
    for s in range(n_streams):
        b = np.roll(bases[s % len(bases)], ((s * 7) % 64, (s * 13) % 64), axis=(0, 1))
        fn[0, s] = b
        for k in range(1, n_sets):   # next frame of the stream: small temporal jitter
            jit = rng.randint(-2, 3, size=(H, W, 1)).astype(np.int16)
            fn[k, s] = np.clip(b.astype(np.int16) + jit, 0, 255).astype(np.uint8)
    boxes = np.stack([synth.make_boxes(n_streams, H, W, rng, lo=96, hi=400) for _ in range(n_sets)])
    if not jpeg:
        return frames, boxes
    import cv2
    streams = []
    for k in range(n_sets):
        row = []
        for s in range(n_streams):
            ok, enc = cv2.imencode(".jpg", fn[k, s], [cv2.IMWRITE_JPEG_QUALITY, 85])
            fn[k, s] = cv2.imdecode(enc, cv2.IMREAD_COLOR)
            row.append(enc.tobytes())
        streams.append(row)
    return frames, boxes, streams


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
def cpu_pipeline(frames, boxes, sd, n_threads, streams=None):
    """The reference's CPU path (oracle port) over frames[t][s]: returns frames processed.  With `streams` the frame is
    first decoded from its JPEG stream with cv2.imdecode, as /analyze does (backend_server.py:140-142)."""
    import cv2
    from oracle import effnet as oeff, faceprep as ofp, forensics as ofor, tracker as otr
    torch.set_num_threads(n_threads)
    cv2.setNumThreads(n_threads)
    n_sets, n_streams = frames.shape[0], frames.shape[1]
    analyzers = [ofor.OracleForensicAnalyzer() for _ in range(n_streams)]
    trackers = [otr.OracleTemporalTracker(detection_threshold=0.55) for _ in range(n_streams)]
    done = 0
    for t in range(n_sets):
        for s in range(n_streams):
            f = frames[t, s] if streams is None else cv2.imdecode(np.frombuffer(streams[t][s], np.uint8), cv2.IMREAD_COLOR)
            r = analyzers[s].analyze(f) if t % 3 == 0 else analyzers[s].analyze_fast(f)
            x = ofp.prepare(f, boxes[t, s])
            p = float(torch.sigmoid(oeff.forward(x, sd)).item())
            p = float(ofp.heuristics(p, boxes[t, s][3], boxes[t, s][2]))
            trackers[s].update(p)
            done += 1
    return done


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle port; the reference's
    third-party CNN package is not installable here), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = 8                                      # frames per step
    sd = synth.make_state_dict()
    frames, boxes, streams = make_inputs(sample, 3, seed=1234, jpeg=True)
    fn = frames.numpy()
    for _ in range(args.warmup):
        cpu_pipeline(fn[:1], boxes[:1], sd, cores, streams[:1])
    t0 = time.perf_counter()
    n = 0
    for k in range(args.steps):
        n += cpu_pipeline(fn[k % 3:k % 3 + 1], boxes[k % 3:k % 3 + 1], sd, cores, streams[k % 3:k % 3 + 1])
    dt = time.perf_counter() - t0
    v = n / dt
    _emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"720p full per-frame path on CPU from JPEG streams (cv2.imdecode + 6 forensic signals + face prep + "
                               f"EfficientNet-B0 fp32 bs=1 + vote), {sample} frames/step, 1 face/frame"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} frames/step x {args.steps} steps of the bench workload"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------
_JSON_OUT = None


def _emit(line):
    """The ONE JSON line goes to the process's original stdout."""
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(line + "\n")
    out.flush()


def _bind_to_gpu_cpus(local):
    """N > 1: run this rank on the CPUs NVML reports as local to its GPU, so the pinned host frames it allocates next land on
    that GPU's NUMA node (four ranks copying from one node share that node's memory and PCIe root).  Best effort: returns the
    CPU list, or None if NVML / the cpuset does not allow it."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[local]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else local
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (max(ncpu, 64) + 63) // 64)
        cpus = {w * 64 + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None


def main():
    # Libraries may print to file descriptor 1 (NCCL prints its version banner there when NCCL_DEBUG=VERSION is set in the
    # environment): keep the real stdout for the JSON line and send everything else written to fd 1 to stderr.
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=6)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=STREAMS)
    ap.add_argument("--dtype", default="fp32", choices=["bf16", "fp32"],
                    help="classifier precision: fp32 = 3xTF32 tensor-core accuracy mode (parity-green, default), bf16 = fastest")
    ap.add_argument("--sustain", type=float, default=5.0, help="seconds of back-to-back steps for the sustained-rate figure (0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graphs", dest="graphs", action="store_false", help="launch every kernel from the host instead of replaying CUDA graphs")
    args = ap.parse_args()
    if os.environ.get("DFD_WATCHDOG"):            # dump the Python stack and exit if the run stalls (debugging aid)
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["DFD_WATCHDOG"]), exit=True)
    if args.impl == "reference":
        return run_reference(args)

    from dfd_b200 import roofline
    from dfd_b200.engine import Engine, RECORD_DTYPE

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    numa = None
    if world > 1:
        numa = _bind_to_gpu_cpus(local)      # before the pinned frame buffers are allocated (first touch decides their NUMA node)
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    S = args.streams
    K, Wm = args.steps, max(args.warmup, 3)
    n_sets = 3

    sd = synth.make_state_dict()
    eng = Engine(device=local, max_streams=S, max_batch=S, max_crop=512, detection_threshold=0.55)
    eng.load_state_dict(sd)
    host_frames, boxes, jpeg_streams = make_inputs(S, n_sets, seed=1234 + rank, jpeg=True)
    jpeg_packed = [eng.pack_jpegs(jpeg_streams[k]) for k in range(n_sets)]     # pinned host buffers + offsets (the wire format)
    dev_frames = host_frames.to(dev)                              # resident inputs for `value`
    dev_boxes = [torch.from_numpy(boxes[k]).to(dev) for k in range(n_sets)]
    sids = torch.arange(S, dtype=torch.int32, device=dev)
    box_frame = torch.arange(S, dtype=torch.int32, device=dev)
    full_flags = [torch.full((S,), int(k == 0), dtype=torch.uint8, device=dev) for k in range(3)]
    # N > 1: streams are sharded by owner(stream) = id mod world (sharding.StreamSharder); this rank's engine numbers them
    # by local slot.  The vote kernel writes its records straight into the sharder's preallocated send buffer.
    sharder = None
    if world > 1:
        from dfd_b200.sharding import StreamSharder
        sharder = StreamSharder(rank, world)
        global_ids = np.arange(world * S)
        own_idx, own_slots = sharder.select(global_ids)
        assert len(own_idx) == S and list(own_slots) == list(range(S))
        rec, gathered = sharder.make_buffers(S, dev)
    else:
        rec = torch.empty(S * RECORD_DTYPE.itemsize, dtype=torch.uint8, device=dev)
        gathered = None
    rec_host = torch.empty(S * RECORD_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # L2 flush buffer (> 126 MB)

    # one CUDA graph per rotating frame set (set k is always analysed with cadence flag k: full / fast / fast)
    graphs = None
    if args.graphs:
        graphs = [eng.capture_step(dev_frames[k], sids, full_flags[k], dev_boxes[k], box_frame, dtype=args.dtype, records_out=rec)
                  for k in range(n_sets)]
        eng.reset(-1)

    def step(i, frames_dev):
        if graphs is not None and frames_dev.data_ptr() == dev_frames[i % n_sets].data_ptr():
            graphs[i % n_sets].replay()
        else:
            eng.analyze_batch(frames_dev, sids, full_flags[i % 3], dev_boxes[i % n_sets], box_frame, dtype=args.dtype,
                              records_out=rec)
        if world > 1:
            sharder.gather_records(rec, S, out=gathered)            # NCCL verdict gather (config 4), no staging copy

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (`value`) ----
    for i in range(Wm):
        step(i, dev_frames[i % n_sets])
    barrier()
    gather_check = None
    if world > 1:
        # every rank must hold every stream's record: world x S distinct global ids, this rank's own segment identical to
        # what its engine wrote, all streams at the same frame count
        allrec = sharder.globalize(gathered, S)
        mine = Engine.records_to_numpy(rec)
        seg = allrec[np.isin(allrec["stream_id"], global_ids[own_idx])]
        ok = (len(allrec) == world * S and sorted(allrec["stream_id"].tolist()) == list(range(world * S))
              and np.array_equal(np.sort(seg["vote_input"]), np.sort(mine["vote_input"]))
              and len(set(allrec["frame_count"].tolist())) == 1)
        gather_check = bool(ok)
        assert ok, "gathered verdict records are inconsistent"
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lpre = eng.launches
    eng.analyze_batch(dev_frames[0], sids, full_flags[0], dev_boxes[0], box_frame, dtype=args.dtype, records_out=rec)
    launches_per_step = eng.launches - lpre          # kernels one step launches (graph replays re-issue the same kernels)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(K):
        step(Wm + i, dev_frames[(Wm + i) % n_sets])               # 707 MB of frames per step >> 126 MB L2
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = launches_per_step * K
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * S * K / (ms / 1e3)

    # ---- end to end through the public call: pinned host frames -> H2D -> path -> D2H records ----
    copy_stream, comp_stream = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    stage = [torch.empty_like(dev_frames[0]) for _ in range(2)]
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def e2e_loop(n, base, jpeg=True):
        for i in range(n):
            b = i % 2
            with torch.cuda.stream(copy_stream):
                if i >= 2:
                    copy_stream.wait_event(consumed[b])
                if jpeg:      # host: header parsing; H2D of the streams; device: unstuff / Huffman / IDCT / colour -> stage[b]
                    pk, off = jpeg_packed[(base + i) % n_sets]
                    eng.decode_jpeg_batch(pk, off, H, W, out=stage[b])
                else:
                    stage[b].copy_(host_frames[(base + i) % n_sets], non_blocking=True)
                copied[b].record(copy_stream)
            with torch.cuda.stream(comp_stream):
                comp_stream.wait_event(copied[b])
                step(base + i, stage[b])
                consumed[b].record(comp_stream)
                rec_host.copy_(rec, non_blocking=True)
        comp_stream.synchronize()

    def timed_e2e(jpeg):
        e2e_loop(3, 0, jpeg)
        barrier()
        with torch.cuda.stream(copy_stream):
            e0.record(copy_stream)
        e2e_loop(K, 3, jpeg)
        with torch.cuda.stream(comp_stream):
            e1.record(comp_stream)
        barrier()
        t_ = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        return t_

    ms_e2e = timed_e2e(True)                 # the /analyze wire format: JPEG streams in, verdict records out
    e2e_value = world * S * K / (float(ms_e2e.item()) / 1e3)
    ms_raw = timed_e2e(False)                # raw decoded frames in (2.76 MB each): PCIe-bound
    e2e_raw_value = world * S * K / (float(ms_raw.item()) / 1e3)
    h2d = int(np.mean([int(off[-1]) for _, off in jpeg_packed]))      # JPEG bytes per step
    h2d_raw = S * H * W * 3
    d2h = S * RECORD_DTYPE.itemsize
    # JPEG decode alone (device time of the ingest stage, streams already in pinned memory)
    dec0, dec1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    dec0.record()
    for i in range(6):
        eng.decode_jpeg_batch(jpeg_packed[i % n_sets][0], jpeg_packed[i % n_sets][1], H, W, out=stage[0])
    dec1.record()
    torch.cuda.synchronize()
    ms_decode = dec0.elapsed_time(dec1) / 6

    # ---- sustained rate: back-to-back steps for >= args.sustain seconds (the timed region above is ~0.1 s) ----
    sustained = None
    if args.sustain > 0:
        barrier()
        s_sampler = ClockSampler(local)
        if rank == 0:
            s_sampler.start()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        est = max(ms / K, 1e-3)
        n_s = max(int(args.sustain * 1e3 / est), K)
        s0.record()
        for i in range(n_s):
            step(i, dev_frames[i % n_sets])
        s1.record()
        barrier()
        t_s = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_s, op=dist.ReduceOp.MAX)
        sc = s_sampler.stop() if rank == 0 else None
        sustained = {"value": world * S * n_s / (float(t_s.item()) / 1e3), "unit": UNIT, "steps": n_s,
                     "seconds": float(t_s.item()) / 1e3, "clocks": sc}

    # ---- per-kernel device times (3 extra steps, one event after every launch) -> roofline ----
    roof, kernels, functions = None, [], []
    if rank == 0:
        torch.cuda.synchronize()
        agg, order = {}, []
        for i in range(3):
            flush.zero_()                                   # cold L2; finished before the first profiled event is recorded
            torch.cuda.synchronize()
            eng.profile_start()
            eng.analyze_batch(dev_frames[i % n_sets], sids, full_flags[i % 3], dev_boxes[i % n_sets], box_frame,
                              dtype=args.dtype, records_out=rec)
            for name, cnt, tms in eng.profile_stop():
                if name not in agg:
                    agg[name] = [0, 0.0]
                    order.append(name)
                agg[name][0] += cnt
                agg[name][1] += tms
        prof = [(name, agg[name][0], agg[name][1]) for name in order]
        hbm, tf, how = peaks()
        area = float(np.mean(boxes[:, :, 2] * boxes[:, :, 3]))
        total_ms = sum(p[2] for p in prof)
        esz = 2 if args.dtype == "bf16" else 4
        for name, cnt, tms in prof:
            by = roofline.kernel_bytes(name, S, H, W, S, area, esz)
            per = tms / cnt
            kernels.append({"kernel": name, "launches": cnt, "ms_per_launch": round(per, 4),
                            "share": round(tms / total_ms, 4),
                            "gbs": round(by / per / 1e6, 1) if by else None})
        kernels.sort(key=lambda k: -k["share"])
        # dominant kernel = the kernel FUNCTION with the largest share of the step (all its launches of one step)
        fn = {}
        for name, cnt, tms in prof:
            f = fn.setdefault(name.split(":")[0], {"ms": 0.0, "bytes": 0, "launches": 0, "unknown": 0})
            by = roofline.kernel_bytes(name, S, H, W, S, area, esz)
            f["ms"] += tms; f["launches"] += cnt
            if by:
                f["bytes"] += by * cnt
            else:
                f["unknown"] += 1
        top_fn = max(fn, key=lambda k: fn[k]["ms"])
        tf_ = fn[top_fn]
        gbs = tf_["bytes"] / tf_["ms"] / 1e6
        traffic = None
        for tname in ("traffic_r02.json", "traffic_r01.json"):          # ncu dram bytes per launch of that kernel (committed)
            tpath = os.path.join(ROOT, "profiles", tname)
            if traffic is None and os.path.exists(tpath):
                with open(tpath) as f:
                    tj = json.load(f)
                ent = tj.get(top_fn) if isinstance(tj.get(top_fn), dict) else (tj if tj.get("kernel") == top_fn else None)
                if ent:
                    traffic = ent.get("dram_bytes_per_launch")
        roof = {"bound": "hbm", "kernel": top_fn, "achieved": round(gbs, 1), "peak": hbm, "unit": "GB/s",
                "frac": round(gbs / hbm, 4), "traffic": traffic, "peak_source": how,
                "launches_per_step": tf_["launches"] // 3,
                "algorithmic_bytes_per_launch": int(tf_["bytes"] / tf_["launches"]),
                "avg_launch_ms": round(tf_["ms"] / tf_["launches"], 5), "share_of_step": round(tf_["ms"] / total_ms, 4),
                "note": "bytes = layer-granular algorithmic bytes summed over this kernel's launches in one step / its summed "
                        "event-timed duration (3 profiled steps, one CUDA event after every launch)"}
        functions = sorted(({"kernel": k, "launches_per_step": v["launches"] // 3, "ms_per_step": round(v["ms"] / 3, 4),
                             "share": round(v["ms"] / total_ms, 4),
                             "gbs": round(v["bytes"] / v["ms"] / 1e6, 1) if v["bytes"] and not v["unknown"] else None}
                            for k, v in fn.items()), key=lambda d: -d["share"])
        try:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", "bench_kernels.json"), "w") as f:
                json.dump({"dtype": args.dtype, "streams": S, "kernels": kernels, "functions": functions, "peak_gbs": hbm}, f, indent=1)
        except OSError:
            pass

    # ---- bs=1 per-frame latency (BASELINE.json metric, second half): one 720p frame + one box, frame resident in
    #      HBM -> verdict record in HBM, CUDA-graph replay; and the same through the host API with H2D / D2H ----
    latency = None
    if rank == 0:
        try:
            one = dev_frames[0][:1].contiguous()
            sid1, full1, bf1 = sids[:1].contiguous(), full_flags[0][:1].contiguous(), box_frame[:1].contiguous()
            box1 = dev_boxes[0][:1].contiguous()
            rec1 = torch.empty(RECORD_DTYPE.itemsize, dtype=torch.uint8, device=dev)
            side = torch.cuda.Stream(dev)
            with torch.cuda.stream(side):
                for _ in range(3):
                    eng.analyze_batch(one, sid1, full1, box1, bf1, dtype=args.dtype, records_out=rec1)
            side.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                eng.analyze_batch(one, sid1, full1, box1, bf1, dtype=args.dtype, records_out=rec1)
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(300)]
            for a, b in evs:
                a.record(); graph.replay(); b.record()
            torch.cuda.synchronize()
            lat = sorted(a.elapsed_time(b) for a, b in evs[50:])
            host1 = host_frames[0][:1]
            t_host = []
            for i in range(120):
                t0 = time.perf_counter()
                one.copy_(host1, non_blocking=True)
                graph.replay()
                rec_host[:RECORD_DTYPE.itemsize].copy_(rec1, non_blocking=True)
                torch.cuda.synchronize()
                t_host.append((time.perf_counter() - t0) * 1e3)
            t_host = sorted(t_host[20:])
            latency = {"bs": 1, "p50_ms": lat[len(lat) // 2], "p99_ms": lat[int(len(lat) * 0.99)],
                       "how": "CUDA-graph replay of dfd_analyze_batch (1 frame 720p, 1 box, full forensics), device-resident, CUDA events",
                       "host_p50_ms": t_host[len(t_host) // 2], "host_p99_ms": t_host[int(len(t_host) * 0.99)],
                       "host_how": "pinned frame -> H2D -> graph replay -> D2H record, wall clock around synchronize"}
        except Exception as e:                                   # never let the extra measurement break the contract line
            latency = {"error": str(e)[:200]}

    # ---- the other BASELINE.json configurations that fit one GPU, as side measurements (rank 0, N=1) ----
    side = None
    if rank == 0 and world == 1:
        try:
            side = {}

            def graph_time(fn, reps):
                st = torch.cuda.Stream(dev)
                with torch.cuda.stream(st):
                    fn(); fn()
                st.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=st):
                    fn()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                for _ in range(3):
                    g.replay()
                a.record()
                for _ in range(reps):
                    flush.zero_()                      # cold L2 between replays (the 256 MB memset is inside the timing: subtracted below)
                    g.replay()
                b.record()
                torch.cuda.synchronize()
                t_all = a.elapsed_time(b)
                a.record()
                for _ in range(reps):
                    flush.zero_()
                b.record()
                torch.cuda.synchronize()
                return (t_all - a.elapsed_time(b)) / reps

            # config 2: EfficientNet-B0 classifier only, 224x224, batch 256 -- bf16 (BASELINE config 2) and the fp32 accuracy mode
            crops = eng.face_prep_batch(dev_frames[0], dev_boxes[0], box_frame, dtype="bf16")
            ms2 = graph_time(lambda: eng.effnet_forward(crops), 10)
            side["config2_classifier_bf16_b256"] = {"crops_per_sec": S / ms2 * 1e3, "ms": ms2,
                                                    "hbm_layer_granular_bound_crops_per_sec": hbm * 1e9 / 27.42e6,
                                                    "parity": "bf16 gate (5e-3) NOT met on the synthetic weights: max |dp| 0.06, "
                                                              "0 / 1920 verdicts and 0.26 % of votes differ (tests/test_gpu_configs.py)"}
            crops32 = eng.face_prep_batch(dev_frames[0], dev_boxes[0], box_frame, dtype="fp32")
            ms2f = graph_time(lambda: eng.effnet_forward(crops32), 10)
            side["config2_classifier_fp32_3xtf32_b256"] = {"crops_per_sec": S / ms2f * 1e3, "ms": ms2f,
                                                           "hbm_layer_granular_bound_crops_per_sec": hbm * 1e9 / 54.84e6,
                                                           "parity": "fp32 gate (1e-4) met: max |dp| 7e-6, verdicts identical"}
            del crops, crops32
            # the whole path in the OTHER precision (device-resident frames, one graph replay per step)
            other = "bf16" if args.dtype == "fp32" else "fp32"
            go = eng.capture_step(dev_frames[0], sids, full_flags[0], dev_boxes[0], box_frame, dtype=other, records_out=rec)
            torch.cuda.synchronize()
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            go.replay()
            a_.record()
            for _ in range(10):
                flush.zero_()
                go.replay()
            b_.record()
            torch.cuda.synchronize()
            t_all = a_.elapsed_time(b_)
            a_.record()
            for _ in range(10):
                flush.zero_()
            b_.record()
            torch.cuda.synchronize()
            mso = (t_all - a_.elapsed_time(b_)) / 10
            side[f"full_path_{other}"] = {"frames_per_sec": S / mso * 1e3, "ms": mso}
            side["jpeg_decode_720p"] = {"frames_per_sec": S / ms_decode * 1e3, "ms_per_step": ms_decode,
                                        "mean_stream_bytes": h2d // S,
                                        "how": "dfd_decode_jpeg_batch alone (host header parsing + H2D of the streams + unstuff / "
                                               "Huffman / IDCT / colour kernels), 256 frames per call"}
            # config 3: six forensic signals on 1080p frames, batch 64 (all frames "full")
            f1080 = torch.randint(0, 256, (64, 1080, 1920, 3), dtype=torch.uint8, device=dev)
            sid64 = torch.arange(64, dtype=torch.int32, device=dev)
            full64 = torch.ones(64, dtype=torch.uint8, device=dev)
            ms3 = graph_time(lambda: eng.forensics_batch(f1080, sid64, full64), 10)
            side["config3_forensics_1080p_b64"] = {"frames_per_sec": 64 / ms3 * 1e3, "ms": ms3}
            del f1080
            # config 5: 4K frames with 8 variable-size face boxes each, fp32 accuracy mode (8 frames = 64 crops per step)
            eng5 = Engine(device=local, max_streams=8, max_batch=64, max_crop=1200, detection_threshold=0.55)
            try:
                eng5.load_state_dict(sd)
                rng5 = np.random.RandomState(77)
                f4k = torch.randint(0, 256, (8, 2160, 3840, 3), dtype=torch.uint8, device=dev)
                bx5 = torch.from_numpy(np.concatenate([synth.make_boxes(8, 2160, 3840, rng5, lo=48, hi=1200) for _ in range(8)])).to(dev)
                bf5 = torch.arange(8, dtype=torch.int32, device=dev).repeat_interleave(8)
                sid5 = torch.arange(8, dtype=torch.int32, device=dev)
                full5 = torch.ones(8, dtype=torch.uint8, device=dev)
                rec5 = torch.empty(8 * RECORD_DTYPE.itemsize, dtype=torch.uint8, device=dev)
                ms5 = graph_time(lambda: eng5.analyze_batch(f4k, sid5, full5, bx5, bf5, dtype="fp32", records_out=rec5), 5)
                side["config5_4k_8boxes_fp32"] = {"frames_per_sec": 8 / ms5 * 1e3, "crops_per_sec": 64 / ms5 * 1e3, "ms": ms5}
            finally:
                eng5.close()
        except Exception as e:
            side = {"error": str(e)[:200]}

    # ---- CPU baseline (rank 0, N=1): the oracle port on a bounded sample of the same workload ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_sample = 48
        fn = host_frames.numpy()[:, :n_sample // 3]
        t0 = time.perf_counter()
        n = cpu_pipeline(fn, boxes[:, :n_sample // 3], sd, cores, [row[:n_sample // 3] for row in jpeg_streams])
        dt = time.perf_counter() - t0
        cpu = {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n} frames (3 consecutive frames of {n_sample // 3} streams) of the bench workload from their JPEG streams "
                         f"(cv2.imdecode + path), fp32 bs=1"}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.dtype == "fp32" else "bf16", "data": "synthetic",
            "config": {"workload": f"720p full per-frame path (6 forensic signals + face prep + EfficientNet-B0 "
                                   f"{'fp32 accuracy mode (3xTF32 tensor cores)' if args.dtype == 'fp32' else 'bf16'} + vote), "
                                   f"{S} streams/GPU, 1 face box/frame, full/fast/fast cadence",
                       "classifier_mode": args.dtype,
                       "parity": ("fp32 gate 1e-4 met (max |dp| 7e-6, verdicts bit-identical over 1920 frames)" if args.dtype == "fp32"
                                  else "bf16 gate 5e-3 NOT met on the synthetic weights (max |dp| 0.06)"),
                       "frames_per_step_per_gpu": S, "frame": "1280x720 BGR u8", "l2": "inputs larger than L2 "
                       "(707 MB of frames per step, 3 rotating sets)", "weights": "fixed-seed random init (synth.make_state_dict)",
                       "submission": "one CUDA graph replay per step" if graphs is not None else "host launches",
                       "collective": "nccl all_gather of 72-B verdict records" if world > 1 else "none"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": float(ms_e2e.item()) / K, "h2d_gbs_per_gpu": round(h2d / (float(ms_e2e.item()) / K) / 1e6, 2),
                    "ingest": "JPEG quality 85 4:2:0 streams (the /analyze wire format) in pinned host memory -> dfd_decode_jpeg_batch "
                              "(host header parsing, H2D, device Huffman / IDCT / colour, bit-exact with cv2.imdecode) -> "
                              "dfd_analyze_batch -> D2H of the 72-B records; ingest of step i+1 overlaps the path of step i",
                    "cpu_affinity": (f"rank 0 bound to the {len(numa)} CPUs local to its GPU" if numa else "unbound")},
            "e2e_raw": {"value": e2e_raw_value, "unit": UNIT, "h2d_bytes_per_step": h2d_raw, "d2h_bytes_per_step": d2h,
                        "ms_per_step": float(ms_raw.item()) / K, "h2d_gbs_per_gpu": round(h2d_raw / (float(ms_raw.item()) / K) / 1e6, 1),
                        "note": "the same with raw decoded frames in pinned host memory (2.76 MB per frame): PCIe-bound"},
            "sustained": sustained, "gather_check": gather_check,
            "gpu_launches": int(launches), "crops_per_sec": value, "clocks": clocks, "roofline": roof,
            "cpu_baseline": cpu, "latency": latency, "other_configs": side, "top_kernels": functions[:6],
        }
        _emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    eng.close()


if __name__ == "__main__":
    main()
