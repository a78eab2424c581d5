#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config: decoded Mpixel/s.

Workload at every N: configs[1] "MSVideo1 16-bit RGB555 1920x1080, 1024-frame batch on 1 B200" per GPU
(weak scaling: every rank decodes its own 1024 independent key frames; the path shards by stream with no
collective).  A step = one decode pass over the whole batch.

  value        whole-job Mpixel/s with bitstreams and pictures resident in HBM (CUDA events, max over ranks)
  e2e          the same through jsp_batch_decode_host with pinned HOST buffers (H2D + decode + D2H timed)
  roofline     msv1_decode kernel: algorithmic bytes / event-timed launch duration vs measured HBM copy peak
  cpu_baseline the CPU oracle (port of the reference decoder) on this box's host cores, bounded sample

`--impl reference` times the reference's CPU algorithm (the oracle port; the Haxe/JS original cannot run
here) on all host cores on a bounded sample of the same workload and prints the same JSON line.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 1920, 1080
MIX = (25, 50, 25)          # % 1-/2-/8-colour blocks -> 8 B per block on average
SEED = 0xC0DEC2
WORKLOAD = "MSVideo1 RGB555 1920x1080 key frames, 25/50/25% 1/2/8-colour blocks, independent streams"


def gen_frames(n, rank, threads=16):
    from jsplayer_b200 import synth
    synth.load()

    def one(i):
        return synth.msv1_frame(False, W, H, SEED + rank * 1000003 + i, mix=MIX)
    with ThreadPoolExecutor(max_workers=threads) as ex:
        return list(ex.map(one, range(n)))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for nm, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture, or None."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("msv1_decode_bytes_per_launch")
        except Exception:
            return None
    return None


def cpu_baseline(frames, threads, budget_s=12.0):
    """The oracle (CPU port of reference src/MSVideo1.hx) on `threads` host threads, bounded sample."""
    import ctypes as C
    from oracle import pyoracle as O
    lib = O.load()
    n = len(frames)
    ln = np.array([len(f) for f in frames], dtype=np.uint32)
    blob = np.frombuffer(b"".join(frames), dtype=np.uint8)
    offs = np.concatenate([[0], np.cumsum(ln.astype(np.uint64))[:-1]]).astype(np.uint64)
    zero = np.zeros(1, dtype=np.uint64)
    key = np.ones(1, dtype=np.uint8)
    descs = (O.StreamDesc * n)()
    for i in range(n):
        d = descs[i]
        d.codec, d.width, d.height, d.bpp = O.CODEC_MSVC16, W, H, 16
        d.palette, d.palette_bytes, d.n_frames = None, 0, 1
        d.bytes = blob.ctypes.data + int(offs[i])
        d.frame_off, d.frame_len, d.frame_key, d.out = zero.ctypes.data, ln[i:].ctypes.data, key.ctypes.data, None
    px = C.c_uint64(0)
    lib.ora_decode_streams_mt(descs, n, threads, 36, C.byref(px))       # warm-up
    times, t_start = [], time.perf_counter()
    while len(times) < 3 or (time.perf_counter() - t_start < budget_s and len(times) < 50):
        times.append(lib.ora_decode_streams_mt(descs, n, threads, 36, C.byref(px)))
    t = statistics.median(times)
    return px.value / t / 1e6, t, len(times)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=1024, help="frames (= independent streams) per GPU")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    warmup = max(3, args.warmup)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return 0
        n_s = max(16, min(args.frames, 2 * cores))
        frames = gen_frames(n_s, 0)
        from oracle import pyoracle as O  # noqa: F401
        per_step = []
        for i in range(warmup + args.steps):
            v, t, reps = cpu_baseline(frames, cores, budget_s=0.0)
            if i >= warmup:
                per_step.append((v, t))
        v = statistics.median([x[0] for x in per_step])
        t = statistics.median([x[1] for x in per_step])
        line = {"impl": "reference", "metric": "decoded Mpixel/s (MSVideo1 RGB555 1080p batch)", "value": v, "unit": "Mpixel/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": WORKLOAD, "frames_per_step": n_s, "width": W, "height": H},
                "cpu_baseline": {"value": v, "unit": "Mpixel/s", "cores": cores, "kind": "port",
                                 "sample": "%d of the workload's frames per step, one frame per thread at a time, median of %d steps" % (n_s, args.steps)},
                "e2e": {"value": v, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from jsplayer_b200 import BatchDecoder, StreamSpec, CodecType, _lib

    frames = gen_frames(args.frames, rank)
    specs = [StreamSpec(CodecType.codec_msvc16, W, H, 16, frames=[f]) for f in frames]
    bd = BatchDecoder(device=local_rank, insignificant_lines=36)
    bd.configure(specs, pinned=True)
    st = bd.stats()
    bd.upload()
    bd.sync()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- parity gate on this rank's data before anything is timed: two frames against the oracle ----
    from oracle import pyoracle as O
    bd.run(); bd.sync()
    outs = [None] * bd.n_frames
    chk = [0, bd.n_frames - 1]
    for i in chk:
        outs[i] = np.empty((H, W), dtype=np.int32)
    bd.download(outs)
    for i in chk:
        exp = O.decode_stream(O.CODEC_MSVC16, W, H, 16, [frames[i]])[0][0]
        if not (outs[i] == exp).all():
            raise SystemExit("bench: GPU output differs from the oracle on frame %d" % i)

    # ---- device-resident timing ----
    sampler = ClockSampler(local_rank)
    barrier()
    bd.time_runs(warmup=warmup, iters=1, flush_l2=False)
    barrier()
    sampler.start()
    ms_total, kms, kcnt = bd.time_runs(warmup=0, iters=args.steps, flush_l2=False)
    barrier()
    clocks = sampler.stop()
    t_local = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_local, op=dist.ReduceOp.MAX)
    ms_total_max = float(t_local.item())
    ms_per_step = ms_total_max / args.steps
    value = st["pixels"] * world / (ms_per_step * 1e-3) / 1e6

    # ---- end to end: pinned host bitstreams -> H2D -> decode -> D2H pinned host pictures ----
    outs_p = bd.alloc_outputs(pinned=True)
    bd.decode_host(outs_p)                       # warm-up (page-touches the pinned output once)
    barrier()
    e2e_t = []
    for _ in range(max(1, args.e2e_steps)):
        t0 = time.perf_counter()
        bd.decode_host(outs_p)
        e2e_t.append(time.perf_counter() - t0)
    barrier()
    e_local = torch.tensor([statistics.mean(e2e_t)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e_local, op=dist.ReduceOp.MAX)
    e2e_s = float(e_local.item())
    e2e_value = st["pixels"] * world / e2e_s / 1e6
    exp = O.decode_stream(O.CODEC_MSVC16, W, H, 16, [frames[1]])[0][0]
    if not (outs_p[1] == exp).all():
        raise SystemExit("bench: end-to-end output differs from the oracle")

    if rank == 0:
        peak, peak_src = hbm_peak()
        n_launch = max(1, kcnt[0])
        k_ms = kms[0] / n_launch                                  # avg msv1_decode launch duration
        achieved = st["alg_bytes"] / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
        traffic = ncu_traffic()
        line = {
            "metric": "decoded Mpixel/s (MSVideo1 RGB555 1080p batch)", "value": value, "unit": "Mpixel/s",
            "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_gpu": args.frames, "width": W, "height": H,
                       "bytes_in_per_gpu": st["in_bytes"], "bytes_out_per_gpu": st["out_bytes"],
                       "l2": "inputs+outputs (%.1f GB) far exceed the 126 MB L2; no flush between steps" % ((st["in_bytes"] + st["out_bytes"]) / 1e9),
                       "parallelism": "stream-sharded x%d, no collective" % world},
            "e2e": {"value": e2e_value, "unit": "Mpixel/s", "h2d_bytes_per_step": st["in_bytes"], "d2h_bytes_per_step": st["out_bytes"],
                    "ms_per_step": e2e_s * 1e3, "steps": len(e2e_t)},
            "gpu_launches": int((kcnt[0] + kcnt[1]) + 2 * args.steps),
            "roofline": {"bound": "hbm", "kernel": "msv1_decode_kernel<false>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "alg_bytes_per_launch": st["alg_bytes"], "launch_ms": k_ms},
            "clocks": clocks,
        }
        if not args.no_cpu_baseline:
            n_s = max(16, min(args.frames, 2 * cores))
            v, t, reps = cpu_baseline(frames[:n_s], cores)
            line["cpu_baseline"] = {"value": v, "unit": "Mpixel/s", "cores": cores, "kind": "port",
                                    "sample": "%d of the %d frames, one frame per thread at a time, median of %d passes (%.2f s each)" % (n_s, args.frames, reps, t)}
        print(json.dumps(line))
    bd.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
