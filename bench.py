#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's configs: decoded Mpixel/s.

Headline workload (every N): configs[1] "MSVideo1 16-bit RGB555 1920x1080, 1024-frame batch on 1 B200" per GPU
(weak scaling: every rank decodes its own 1024 independent key frames; the path shards by stream with no
collective).  A step = one decode pass over the whole batch.

  value        whole-job Mpixel/s with bitstreams and pictures resident in HBM (CUDA events, max over ranks)
  e2e          the same through jsp_batch_decode_host with pinned HOST buffers (H2D + decode + D2H timed)
  roofline     dominant kernel: algorithmic bytes / event-timed launch duration vs measured HBM copy peak
  cpu_baseline the CPU oracle (port of the reference decoder) on this box's host cores, bounded sample
  codecs       (rank 0, N = 1 only, --no-codecs to skip) the same measurement for the other BASELINE configs at their
               stated sizes: configs[0] (c1: MSVideo1 8-bit 320x240, 300 frames), configs[2] (c3: ScreenPressor 1280x720
               keyframe-only, 256 streams), configs[3] (c4: ScreenPressor 1080p, 512 streams of 1 I + 31 P), configs[4]
               (c5: the mixed 4K AVI corpus, this GPU's 8 of the 64 files) and two MSVideo1 inter-frame legs (c2p / c2p8:
               RGB555 and 8-bit 1080p, key + 15 P frames with 85 % skipped blocks)

`--workload c1|c2p|c2p8|c3|c4|c5` makes another config the timed workload instead (same JSON contract).
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

INSIGN = 36                                       # Manager.hx:61 insignificant_lines


# ------------------------------------------------------------------------------------------ workloads ----
class Workload:
    """A list of independent streams (StreamSpec) per rank + how to describe and sample it."""
    name = ""
    metric = ""
    desc = ""
    dominant = 0                                  # JSP_K_* class whose launches the roofline is quoted on
    dominant_name = ""

    def specs(self, rank, scale):                 # -> list of StreamSpec
        raise NotImplementedError


class C2(Workload):
    name = "c2"
    metric = "decoded Mpixel/s (MSVideo1 RGB555 1080p batch)"
    W, H = 1920, 1080
    MIX = (25, 50, 25)                            # % 1-/2-/8-colour blocks -> 8 B per block
    SEED = 0xC0DEC2
    desc = "MSVideo1 RGB555 1920x1080 key frames, 25/50/25% 1/2/8-colour blocks, independent streams"
    dominant, dominant_name = 0, "msv1_decode_kernel<false,false>"

    def __init__(self, frames=1024, mix=None):
        self.n = frames
        if mix:
            self.MIX = tuple(mix)
            self.desc = "MSVideo1 RGB555 1920x1080 key frames, %d/%d/%d%% 1/2/8-colour blocks (sweep), independent streams" % self.MIX

    def frames(self, rank, n=None, threads=16):
        import synth
        synth.load()
        n = self.n if n is None else n

        def one(i):
            return synth.msv1_frame(False, self.W, self.H, self.SEED + rank * 1000003 + i, mix=self.MIX)
        with ThreadPoolExecutor(max_workers=threads) as ex:
            return list(ex.map(one, range(n)))

    def specs(self, rank, n=None):
        from jsplayer_b200 import StreamSpec, CodecType
        return [StreamSpec(CodecType.codec_msvc16, self.W, self.H, 16, frames=[f]) for f in self.frames(rank, n)]

    def config(self):
        return {"workload": self.desc, "frames_per_gpu": self.n, "width": self.W, "height": self.H}


class MSV1Streams(Workload):
    """MSVideo1 streams with P frames: a key frame, then P frames whose blocks are skipped in runs (SURVEY.md 8d C1 recipe:
    85 % skipped, geometric runs of mean 40 carried across rows, coded blocks 40/40/20 % 1-/2-/8-colour)."""
    dominant = -1                                 # whichever kernel takes most of the step (the wide copy of sparse P frames)

    def __init__(self, name, is8, width, height, streams, frames_per_stream, seed, metric, desc):
        self.name, self.is8, self.W, self.H, self.n, self.fps, self.seed = name, is8, width, height, streams, frames_per_stream, seed
        self.metric, self.desc = metric, desc
        self.dominant_name = "msv1_decode_kernel<%s,false>" % ("true" if is8 else "false")

    def specs(self, rank, n=None):
        from jsplayer_b200 import StreamSpec, CodecType
        import synth
        synth.load()
        n = self.n if n is None else n
        pal = synth.random_palette(self.seed) if self.is8 else None

        def one(i):
            sd = self.seed + rank * 1000003 + i * 4099
            fr = [synth.msv1_frame(self.is8, self.W, self.H, sd, mix=(40, 40, 20))]
            fr += [synth.msv1_frame(self.is8, self.W, self.H, sd + f, skip_permille=850, mean_skip=40, mix=(40, 40, 20)) for f in range(1, self.fps)]
            return fr
        with ThreadPoolExecutor(max_workers=16) as ex:
            streams = list(ex.map(one, range(n)))
        return [StreamSpec(CodecType.codec_msvc8 if self.is8 else CodecType.codec_msvc16, self.W, self.H, 8 if self.is8 else 16,
                           frames=fr, palette=pal) for fr in streams]

    def config(self):
        return {"workload": self.desc, "streams_per_gpu": self.n, "frames_per_stream": self.fps, "width": self.W, "height": self.H}


class SPWorkload(Workload):
    """ScreenPressor streams from the screen-content generator.  `distinct` different streams are encoded and
    repeated round-robin up to `streams` (every copy is decoded independently with its own model state)."""
    dominant, dominant_name = 2, "sp_decode_kernel"

    def __init__(self, streams, frames_per_stream, width, height, gop, versions, distinct, seed, change_permille, name, metric, desc):
        self.n, self.fps, self.W, self.H, self.gop = streams, frames_per_stream, width, height, gop
        self.versions, self.distinct, self.seed, self.change = versions, distinct, seed, change_permille
        self.name, self.metric, self.desc = name, metric, desc

    def specs(self, rank, n=None):
        from jsplayer_b200 import StreamSpec, CodecType
        import synth
        synth.load()
        n = self.n if n is None else n
        k = min(self.distinct, n)

        def one(i):
            ver = self.versions[i % len(self.versions)]
            return synth.sp_stream(self.W, self.H, self.fps, seed=self.seed + rank * 7919 + i, version=ver,
                                   gop=self.gop, change_permille=self.change)[:2]
        with ThreadPoolExecutor(max_workers=16) as ex:
            base = list(ex.map(one, range(k)))
        return [StreamSpec(CodecType.codec_screenpressor, self.W, self.H, 24, frames=base[i % k][0], keys=base[i % k][1])
                for i in range(n)]

    def config(self):
        return {"workload": self.desc, "streams_per_gpu": self.n, "frames_per_stream": self.fps, "width": self.W,
                "height": self.H, "distinct_streams": min(self.distinct, self.n),
                "stream_versions": ["v%d" % v for v in self.versions]}


class C5(Workload):
    """BASELINE.json configs[4]: mixed 4K corpus of real RIFF AVI files (MSVideo1 RGB555 + ScreenPressor v2/v3/v4, key
    frame every 16), written to disk, read back through the AVI indexer into pinned buffers and decoded as
    keyframe-delimited segments (GOPs) -- the unit of sharding; 64 files over 8 GPUs = 8 files per GPU."""
    name = "c5"
    metric = "decoded Mpixel/s (mixed 4K AVI corpus, GOP-sharded)"
    W, H, FRAMES, GOP = 3840, 2160, 64, 16
    desc = "mixed 3840x2160 corpus: MSVideo1 RGB555 + ScreenPressor AVI files, 64 frames, key every 16, via the AVI indexer, GOP-sharded"
    dominant, dominant_name = 0, "msv1_decode_kernel<false,false>"

    def __init__(self, files):
        self.n = files
        self._keep = []

    def specs(self, rank, n=None):
        import tempfile
        from jsplayer_b200 import avi
        import synth
        from synth.avi import write_avi
        synth.load()
        n = self.n if n is None else n
        tmp = tempfile.mkdtemp(prefix="jsp_c5_")
        W, H = self.W, self.H

        def one(i):
            path = os.path.join(tmp, "r%d_f%d.avi" % (rank, i))
            seed = 0xC0DEC5 + rank * 7919 + i
            if i % 2 == 0:
                frames = [synth.msv1_frame(False, W, H, seed * 64 + f, skip_permille=0 if f % self.GOP == 0 else 850)
                          for f in range(self.FRAMES)]
                keys = [1 if f % self.GOP == 0 else 0 for f in range(self.FRAMES)]
                write_avi(path, W, H, 16, b"CRAM", frames, keys)
            else:
                frames, keys, _ = synth.sp_stream(W, H, self.FRAMES, seed=seed, version=2 + (i // 2) % 3, gop=self.GOP, change_permille=20)
                write_avi(path, W, H, 24, b"SCPR", frames, keys)
            return path
        with ThreadPoolExecutor(max_workers=8) as ex:
            paths = list(ex.map(one, range(n)))
        from jsplayer_b200 import _lib
        pinned = _lib.load().jsp_device_count() > 0          # (the reference arm also runs on a box without a GPU)
        streams = [avi.load_avi(p, pinned=pinned) for p in paths]
        for p in paths:
            os.unlink(p)
        os.rmdir(tmp)
        self._keep = streams
        specs, self.where = avi.gop_specs(streams)
        return specs

    def config(self):
        return {"workload": self.desc, "files_per_gpu": self.n, "frames_per_file": self.FRAMES, "gop": self.GOP,
                "width": self.W, "height": self.H}


def spec_frames(sp):
    """The compressed frames of a StreamSpec as a list of bytes."""
    if sp.bytes_buf is None:
        return [bytes(f) for f in sp.frames]
    return [sp.bytes_buf[int(o):int(o) + int(n)].tobytes() for o, n in zip(sp.frame_off, sp.frame_len)]


def make_workload(name, args):
    if name == "c5":
        return C5(args.files)
    if name == "c2":
        return C2(args.frames, args.c2_mix)
    if name == "c1":
        return MSV1Streams("c1", True, 320, 240, args.streams or 1, 300, 0xC0DEC1,
                           "decoded Mpixel/s (MSVideo1 8-bit 320x240, 300 frames)",
                           "MSVideo1 8-bit palettised 320x240, 300 frames (key + 299 P, 85 % of blocks skipped): BASELINE configs[0], the reference's own CPU case -- ONE stream, so the GPU runs 300 dependent launches")
    if name in ("c2p", "c2p8"):
        is8 = name == "c2p8"
        return MSV1Streams(name, is8, 1920, 1080, args.streams or 64, 16, 0xC0DE2B if is8 else 0xC0DE2A,
                           "decoded Mpixel/s (MSVideo1 %s 1080p inter-frame streams)" % ("8-bit" if is8 else "RGB555"),
                           "MSVideo1 %s 1920x1080 streams of 1 key + 15 P frames, 85 %% of the P frames' blocks skipped (copied from the previous picture)" % ("8-bit palettised" if is8 else "RGB555"))
    if name == "c3":
        return SPWorkload(args.streams or 256, 4, 1280, 720, 1, args.sp_versions, 32, 0xC0DEC3, 40, "c3",
                          "decoded Mpixel/s (ScreenPressor RGB24 720p key frames)",
                          "ScreenPressor RGB24 1280x720 keyframe-only, 4 I frames per stream, independent synthetic screen-content streams")
    if name == "c4":
        return SPWorkload(args.streams or 512, 32, 1920, 1080, 0, args.sp_versions, 16, 0xC0DEC4, 20, "c4",
                          "decoded Mpixel/s (ScreenPressor 1080p inter-frame streams)",
                          "ScreenPressor 1920x1080 inter-frame streams (1 I + 31 P, skip/copy-heavy screen content), per-stream entropy decode + wide copy")
    raise SystemExit("unknown workload " + name)


# ------------------------------------------------------------------------------------------ helpers ----
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []          # (arrival time, text)
        self.t_mark = None

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
            self.lines.append((time.perf_counter(), ln.strip()))

    def mark(self):
        """The timed region starts now: only samples that arrive from here on count."""
        self.t_mark = time.perf_counter()

    def count(self):
        return sum(1 for t, _ in self.lines if self.t_mark is None or t >= self.t_mark)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, ln in self.lines:
            if self.t_mark is not None and t < self.t_mark:
                continue
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


def ncu_traffic(kernel_key):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture, or None."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(kernel_key)
        except Exception:
            return None
    return None


def oracle_descs(specs):
    """StreamSpec list -> (ctypes array of the oracle's stream descriptors, keep-alive list)."""
    from oracle import pyoracle as O
    keep = []
    descs = (O.StreamDesc * len(specs))()
    cache = {}
    for i, sp in enumerate(specs):
        key = id(sp.frames) if sp.bytes_buf is None else (id(sp.bytes_buf), int(sp.frame_off[0]) if len(sp.frame_off) else 0)
        if key not in cache:
            frames = spec_frames(sp)
            ln = np.array([len(f) for f in frames], dtype=np.uint32)
            off = np.zeros(len(ln), dtype=np.uint64)
            if len(ln):
                off[1:] = np.cumsum(ln.astype(np.uint64))[:-1]
            blob = np.frombuffer(b"".join(frames) + b"\0", dtype=np.uint8).copy()
            keys = np.zeros(len(ln), dtype=np.uint8)
            if sp.keys is None:
                keys[:1] = 1
            else:
                keys[:] = np.asarray(sp.keys, dtype=np.uint8)
            cache[key] = (blob, off, ln, keys)
            keep.append(cache[key])
        blob, off, ln, keys = cache[key]
        d = descs[i]
        d.codec, d.width, d.height, d.bpp = int(sp.codec), sp.width, sp.height, sp.bpp
        pal = np.frombuffer(sp.palette, dtype=np.uint8).copy() if sp.palette else None
        keep.append(pal)
        d.palette, d.palette_bytes, d.n_frames = (pal.ctypes.data if pal is not None else None), (pal.size if pal is not None else 0), len(ln)
        d.bytes, d.frame_off, d.frame_len, d.frame_key, d.out = blob.ctypes.data, off.ctypes.data, ln.ctypes.data, keys.ctypes.data, None
    return descs, keep


def cpu_baseline(specs, threads, budget_s=12.0, min_reps=3):
    """The oracle (CPU port of the reference decoders) on `threads` host threads over `specs`, bounded sample.
    Returns (Mpixel/s, seconds per pass, passes)."""
    import ctypes as C
    from oracle import pyoracle as O
    lib = O.load()
    descs, keep = oracle_descs(specs)
    px = C.c_uint64(0)
    lib.ora_decode_streams_mt(descs, len(specs), threads, INSIGN, C.byref(px))       # warm-up
    times, t_start = [], time.perf_counter()
    while len(times) < min_reps or (time.perf_counter() - t_start < budget_s and len(times) < 5000):
        times.append(lib.ora_decode_streams_mt(descs, len(specs), threads, INSIGN, C.byref(px)))
    t = statistics.median(times)
    return px.value / t / 1e6, t, len(times)


def check_against_oracle(bd, specs, outs, which):
    """Parity gate: pictures of the streams in `which` (all frames) against the CPU oracle."""
    from oracle import pyoracle as O
    first = 0
    firsts = []
    for sp in specs:
        firsts.append(first)
        first += sp.n_frames
    for s in which:
        sp = specs[s]
        exp = O.decode_stream(int(sp.codec), sp.width, sp.height, sp.bpp, spec_frames(sp), keys=sp.keys, palette=sp.palette,
                              insignificant_lines=INSIGN)[0]
        for f in range(sp.n_frames):
            if not (outs[firsts[s] + f].reshape(sp.height, sp.width) == exp[f]).all():
                raise SystemExit("bench: GPU output differs from the oracle (stream %d frame %d)" % (s, f))


def sample_size(wl, cores, n_specs):
    if wl.name == "c2":
        return max(16, min(n_specs, 2 * cores))
    if wl.name in ("c1", "c5"):                     # c5: n_specs counts AVI files (reference arm) or their GOP segments (cpu_baseline)
        return max(1, min(n_specs, cores if wl.name == "c1" else 64))
    return max(8, min(n_specs, cores))


def committed_json(name):
    """A small JSON file under profiles/ (numbers read out of committed ncu captures / box measurements), or {}."""
    p = os.path.join(ROOT, "profiles", name)
    try:
        return json.load(open(p))
    except Exception:
        return {}


RING_ABOVE = 24 << 30          # end-to-end pictures beyond this go through a bounded pinned ring (what a player's buffer ring does)
RING_BYTES = 8 << 30


def measure(wl, args, rank, local_rank, world, dist, torch, with_e2e=True, with_cpu=True, cpu_budget=None):
    """Times `wl` on this rank; returns the JSON line (dict) on rank 0, else None."""
    from jsplayer_b200 import BatchDecoder
    warmup = max(3, args.warmup)
    cores = os.cpu_count() or 1
    specs = wl.specs(rank)
    # one process per GPU: this rank's thread and the pinned buffers it allocates move to the GPU's NUMA node
    bd = BatchDecoder(device=local_rank, insignificant_lines=INSIGN, numa_bind=not args.no_numa_bind)
    bd.configure(specs, pinned=True)
    st = bd.stats()
    kbytes = bd.kernel_bytes()
    bd.upload()
    bd.sync()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- parity gate on this rank's data before anything is timed: four streams spread over the batch ----
    bd.run(); bd.sync()
    first_of = np.cumsum([0] + [sp.n_frames for sp in specs])
    n_sp = len(specs)
    chk = sorted({0, n_sp // 3, (2 * n_sp) // 3, n_sp - 1})
    outs = [None] * bd.n_frames
    for s in chk:
        for f in range(specs[s].n_frames):
            outs[first_of[s] + f] = np.empty((specs[s].height, specs[s].width), dtype=np.int32)
    _, flags = bd.download(outs)
    if (flags & 4).any():
        raise SystemExit("bench: %d frames reported a decode error" % int((flags & 4).astype(bool).sum()))
    check_against_oracle(bd, specs, outs, chk)
    del outs

    # ---- device-resident timing ----
    flush = st["in_bytes"] + st["out_bytes"] < (512 << 20)          # small working sets: flush the 126 MB L2 between steps
    sampler = ClockSampler(local_rank)
    sampler.start()                      # nvidia-smi needs up to a second to produce its first line (longer with 8 ranks)
    barrier()
    bd.time_runs(warmup=warmup, iters=1, flush_l2=flush)
    barrier()
    sampler.mark()
    ms_total, kms, kcnt = bd.time_runs(warmup=0, iters=args.steps, flush_l2=flush)
    # the timed region of a fast workload is shorter than nvidia-smi's 100 ms period: keep the same load running
    # (untimed) until two samples have been taken under it
    t_wait = time.perf_counter()
    while sampler.proc and sampler.count() < 2 and time.perf_counter() - t_wait < 8.0:
        bd.time_runs(warmup=0, iters=args.steps, flush_l2=flush)
    barrier()
    clocks = sampler.stop()
    n_symbols = bd.symbols()
    t_local = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_local, op=dist.ReduceOp.MAX)
    ms_per_step = float(t_local.item()) / args.steps
    value = st["pixels"] * world / (ms_per_step * 1e-3) / 1e6

    # ---- fused display store (JSP_BATCH_DISPLAY): the same batch with the MSVideo1 kernel storing canvas
    #      words bottom-up, checked against the oracle's decode + display conversion, timed like `value` ----
    display_fused = None
    if world == 1 and wl.name in ("c2", "c2p", "c2p8") and not getattr(args, "no_display_leg", False):
        from oracle import pyoracle as O
        bdd = BatchDecoder(device=local_rank, insignificant_lines=INSIGN, display=True, display_flip=True)
        bdd.configure(specs, pinned=True)
        bdd.upload(); bdd.run(); bdd.sync()
        sp = specs[chk[-1]]
        o = [None] * bdd.n_frames
        for f in range(sp.n_frames):
            o[first_of[chk[-1]] + f] = np.empty((sp.height, sp.width), dtype=np.int32)
        bdd.download(o)
        exp = O.decode_stream(int(sp.codec), sp.width, sp.height, sp.bpp, spec_frames(sp), keys=sp.keys, palette=sp.palette,
                              insignificant_lines=INSIGN)[0]
        for f in range(sp.n_frames):
            if not (o[first_of[chk[-1]] + f] == O.display_convert(exp[f], flip=True)).all():
                raise SystemExit("bench: fused display store differs from the oracle (stream %d frame %d)" % (chk[-1], f))
        del o
        bdd.time_runs(warmup=warmup, iters=1, flush_l2=flush)
        torch.cuda.synchronize()
        ms_d, _, _ = bdd.time_runs(warmup=0, iters=args.steps, flush_l2=flush)
        bdd.close()
        display_fused = {"value": st["pixels"] / (ms_d / args.steps * 1e-3) / 1e6, "unit": "Mpixel/s", "ms_per_step": ms_d / args.steps,
                         "ms_per_step_plain": ms_per_step,
                         "what": "device-resident decode with JSP_BATCH_DISPLAY | JSP_BATCH_DISPLAY_FLIP: pictures leave the decode kernel as "
                                 "canvas R,G,B,A words, bottom-up (Manager.hx:363-381, Main.hx:946); 0 extra bytes per pixel against the "
                                 "8 B/pixel of the separate pass; stream %d checked against the oracle's display conversion" % chk[-1]}

    # ---- end to end: pinned host bitstreams -> H2D -> decode -> D2H pinned host pictures ----
    e2e = None
    e2e_steps = args.steps if args.e2e_steps < 0 else args.e2e_steps
    if with_e2e and e2e_steps > 0:
        ring = st["out_bytes"] > RING_ABOVE
        outs_p = bd.alloc_outputs(pinned=True, ring_bytes=RING_BYTES if ring else 0, keep_streams=chk[-1:])
        bd.decode_host(outs_p)                       # warm-up (page-touches the pinned output once)
        barrier()
        e2e_t = []
        for _ in range(e2e_steps):
            t0 = time.perf_counter()
            bd.decode_host(outs_p)
            e2e_t.append(time.perf_counter() - t0)
        barrier()
        e_local = torch.tensor([statistics.mean(e2e_t)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(e_local, op=dist.ReduceOp.MAX)
        e2e_s = float(e_local.item())
        check_against_oracle(bd, specs, outs_p, chk[-1:])
        e2e = {"value": st["pixels"] * world / e2e_s / 1e6, "unit": "Mpixel/s", "h2d_bytes_per_step": st["in_bytes"],
               "d2h_bytes_per_step": st["out_bytes"], "ms_per_step": e2e_s * 1e3, "steps": len(e2e_t),
               "d2h_gbs_all_ranks": st["out_bytes"] * world / e2e_s / 1e9,
               "host_buffers": ("pinned ring of %.0f GB (pictures recycled as a player's buffer ring does; every byte still crosses PCIe)" % (RING_BYTES / 2**30))
               if ring else "one pinned picture per frame"}
        # What THIS box's host takes: all ranks at once do nothing but device -> pinned-host copies (1 GiB x 4 each).  Boxes of the
        # pool differ (two GPUs reached 70 GB/s together on one box and 107 on another; profiles/r02_d2h_ceiling.txt holds one
        # 8-GPU box: 55.8 / 69.8 / 71.1 / 92.0 GB/s at N = 1 / 2 / 4 / 8), so the ceiling is measured live beside the number.
        from jsplayer_b200 import _lib as _L
        barrier()
        mine = _L.load().jsp_host_d2h_gbs(local_rank, 1 << 30, 4)
        c_local = torch.tensor([mine], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(c_local, op=dist.ReduceOp.SUM)
        e2e["host_ceiling_gbs"] = float(c_local.item())
        e2e["host_ceiling_how"] = "all %d ranks at once: 4 device->pinned-host copies of 1 GiB each, summed" % world

    # ---- opt-in in-place end-to-end path (inter-frame workloads): one pinned picture per STREAM, only changed blocks cross PCIe ----
    e2e_inplace = None
    if with_e2e and e2e_steps > 0 and max(sp.n_frames for sp in specs) > 1:
        pics = bd.alloc_stream_pictures(pinned=True)
        s_chk = chk[-1]
        seen = {}

        def on_frame(stream, frame, picture, fl):           # keep every frame of ONE stream to check it afterwards
            if stream == s_chk:
                seen[frame] = picture.copy()
        bd.decode_host_delta(pics, on_frame)                 # warm-up (staging buffers, page touches) + parity material
        sp = specs[s_chk]
        from oracle import pyoracle as O
        exp = O.decode_stream(int(sp.codec), sp.width, sp.height, sp.bpp, spec_frames(sp), keys=sp.keys, palette=sp.palette, insignificant_lines=INSIGN)[0]
        bw, bh = (sp.width, sp.height) if int(sp.codec) == 0 else (sp.width & ~3, sp.height & ~3)
        for f in range(sp.n_frames):
            if f not in seen or not (seen[f][:bh, :bw] == exp[f][:bh, :bw]).all():
                raise SystemExit("bench: in-place path differs from the oracle (stream %d frame %d)" % (s_chk, f))
        barrier()
        ti = []
        for _ in range(e2e_steps):
            t0 = time.perf_counter()
            bd.decode_host_delta(pics, None)
            ti.append(time.perf_counter() - t0)
        barrier()
        i_local = torch.tensor([statistics.mean(ti)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(i_local, op=dist.ReduceOp.MAX)
        i_s = float(i_local.item())
        e2e_inplace = {"value": st["pixels"] * world / i_s / 1e6, "unit": "Mpixel/s", "ms_per_step": i_s * 1e3, "steps": len(ti),
                       "h2d_bytes_per_step": st["in_bytes"], "d2h_bytes_per_step": bd.delta_bytes(),
                       "contract": "jsp_batch_decode_host_delta (opt-in): one host picture per stream updated in place, every frame visible "
                                   "in order through a callback; only 16x16 blocks that differ from the previous picture cross PCIe"}

    line = None
    if rank == 0:
        from jsplayer_b200 import _lib
        peak, peak_src = hbm_peak()
        k = wl.dominant
        if k < 0 or kcnt[k] == 0:                     # e.g. a mixed-coder ScreenPressor run, or inter-frame MSVideo1 (wide copy + decode)
            k = max(range(len(kcnt)), key=lambda i: kms[i])
        n_launch = max(1, kcnt[k])
        k_ms = kms[k] / n_launch                      # average launch duration of the dominant kernel
        share = kms[k] / max(1e-9, sum(kms))
        # algorithmic bytes of ONE (average) launch of the dominant kernel (DESIGN.md "Algorithmic bytes")
        alg_launch = kbytes[k] * args.steps / n_launch
        achieved = alg_launch / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
        cfg = wl.config()
        cfg.update({"bytes_in_per_gpu": st["in_bytes"], "bytes_out_per_gpu": st["out_bytes"],
                    "l2": ("inputs+outputs (%.1f GB) far exceed the 126 MB L2; no flush between steps" % ((st["in_bytes"] + st["out_bytes"]) / 1e9))
                    if not flush else "512 MB memset between steps flushes the 126 MB L2",
                    "parallelism": "stream-sharded x%d, no collective" % world,
                    "parity_gate": "streams %s of this rank against the CPU oracle before timing, the last one again after the end-to-end passes" % chk})
        line = {
            "metric": wl.metric, "value": value, "unit": "Mpixel/s",
            "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": cfg,
            "gpu_launches": int(sum(kcnt) + 2 * args.steps),
            "kernels": {_lib.KERNEL_NAMES[i]: {"launches": int(kcnt[i]), "ms": round(kms[i], 4)} for i in range(len(kcnt)) if kcnt[i]},
            "clocks": clocks,
        }
        hbm = {"bound": "hbm", "kernel": wl.dominant_name if (wl.dominant_name.startswith("msv1") and k == 0) else _lib.KERNEL_NAMES[k],
               "achieved": achieved, "peak": peak, "unit": "GB/s",
               "frac": achieved / peak, "traffic": ncu_traffic(_lib.KERNEL_NAMES[k] + "_bytes_per_launch"),
               "peak_source": peak_src, "alg_bytes_per_launch": alg_launch, "launch_ms": k_ms,
               "share_of_step": share}
        if n_symbols:
            # The entropy stage is one dependent instruction chain per stream: bound by latency, not by bytes or issue slots.
            # Its unit is symbols / s and cycles per symbol per stream; the counters that justify "latency" (issue slots, warp
            # slots and the shared-memory pipe all nearly idle) come from the committed ncu capture of the same kernels.
            ent = (2, 3, 6)
            ent_ms = sum(kms[i] for i in ent) / args.steps
            ent_launches = sum(kcnt[i] for i in ent) / args.steps
            jobs_per_launch = bd.n_frames / max(1.0, ent_launches)
            mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0
            line["entropy"] = {"symbols_per_step": n_symbols, "symbols_per_pixel": n_symbols / st["pixels"],
                               "msymbols_per_s": n_symbols / (ms_per_step * 1e-3) / 1e6,
                               "entropy_kernel_ms_per_step": ent_ms}
            lat = {"bound": "latency", "kernel": "sp2_{rc,ans}_{i,p}_kernel (one warp pair per independent segment)",
                   "achieved": n_symbols / (ms_per_step * 1e-3) / 1e6, "unit": "Msymbol/s", "peak": None, "frac": None,
                   "cycles_per_symbol": ent_ms * 1e-3 * mhz * 1e6 * jobs_per_launch / n_symbols,
                   "jobs_per_launch": jobs_per_launch, "share_of_step": sum(kms[i] for i in ent) / max(1e-9, sum(kms)),
                   "hbm_frac_for_reference": achieved / peak}
            lat.update(committed_json("r02_entropy_counters.json"))
            line["roofline"] = lat
        else:
            if hbm["frac"] < 0.02:
                # a handful of CTAs per launch and hundreds of dependent launches (one stream: every P frame waits for its
                # predecessor): the step is bound by launch latency, not by bytes -- say so instead of a meaningless fraction
                hbm["bound_note"] = "launch-latency bound: %d dependent launches of %.1f us each per step" % (
                    int(sum(kcnt) / args.steps), 1e3 * ms_per_step / max(1.0, sum(kcnt) / args.steps))
            line["roofline"] = hbm
        line["e2e"] = e2e if e2e else {"value": None, "unit": "Mpixel/s", "skipped": "--e2e-steps 0"}
        if e2e_inplace:
            line["e2e_inplace"] = e2e_inplace
        if display_fused:
            line["display_fused"] = display_fused
        if with_cpu and not args.no_cpu_baseline:
            n_s = sample_size(wl, cores, len(specs))
            v, t, reps = cpu_baseline(specs[:n_s], cores, budget_s=cpu_budget or args.cpu_budget)
            line["cpu_baseline"] = {"value": v, "unit": "Mpixel/s", "cores": min(cores, len(specs[:n_s])), "kind": "port",
                                    "sample": "%d of the %d streams, one stream per thread at a time, pictures into a 2-buffer scratch ring per thread; median of %d passes (%.3f s each, ~%.0f s of CPU wall time)" % (n_s, len(specs), reps, t, reps * t)}
    bd.close()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c2p", "c2p8", "c3", "c4", "c5"])
    ap.add_argument("--files", type=int, default=8, help="c5: AVI files per GPU")
    ap.add_argument("--c2-mix", type=lambda s: [int(x) for x in s.split(",")], default=None,
                    help="c2 sweep: percentages of 1-/2-/8-colour blocks, e.g. 100,0,0 (default 25,50,25 = the quoted config)")
    ap.add_argument("--frames", type=int, default=1024, help="c2: frames (= independent streams) per GPU")
    ap.add_argument("--streams", type=int, default=0, help="streams per GPU (defaults: c1 1, c2p/c2p8 64, c3 256, c4 512)")
    ap.add_argument("--sp-versions", type=lambda s: [int(x) for x in s.split(",")], default=[2, 4],
                    help="ScreenPressor stream versions to mix (2 = range coder, 3/4 = rANS)")
    ap.add_argument("--e2e-steps", type=int, default=-1, help="end-to-end passes (default: --steps; 0 skips the leg)")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of wall time for the cpu_baseline sample")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not move this rank's thread / pinned memory to its GPU's NUMA node")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-codecs", action="store_true", help="skip the per-codec ScreenPressor legs")
    ap.add_argument("--no-display-leg", action="store_true", help="skip the fused-display-store leg of the MSVideo1 workloads")
    args = ap.parse_args()
    warmup = max(3, args.warmup)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    wl = make_workload(args.workload, args)

    if args.impl == "reference":
        if rank != 0:
            return 0
        n_s = sample_size(wl, cores, getattr(wl, "n", 1 << 30))
        specs = wl.specs(0, n_s)
        per_step = []
        for i in range(warmup + args.steps):
            # one step = the bounded sample decoded repeatedly for about a second (median pass time reported)
            v, t, reps = cpu_baseline(specs, cores, budget_s=1.0 if i >= warmup else 0.2, min_reps=1)
            if i >= warmup:
                per_step.append((v, t))
        v = statistics.median([x[0] for x in per_step])
        t = statistics.median([x[1] for x in per_step])
        cfg = wl.config()
        cfg["streams_per_step"] = n_s
        n_full = getattr(wl, "n", n_s)
        cfg["sample"] = "%d/%d" % (n_s, n_full)       # a step decodes this many of the workload's streams (bounded CPU time)
        cfg["out"] = "scratch ring (2 pictures per thread): the CPU arm keeps no pictures, which favours it"
        cfg["same_config"] = n_s == n_full
        line = {"impl": "reference", "metric": wl.metric, "value": v, "unit": "Mpixel/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": cfg,
                "cpu_baseline": {"value": v, "unit": "Mpixel/s", "cores": min(cores, len(specs)), "kind": "port",
                                 "sample": "%d of the workload's streams, one stream per thread at a time; a step repeats the sample for ~1 s (median pass), median of %d steps" % (n_s, args.steps)},
                "e2e": {"value": v, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    line = measure(wl, args, rank, local_rank, world, dist, torch)

    # ---- per-codec legs (BASELINE.json's metric is "per codec"), rank 0 of a 1-GPU run only: every other BASELINE config
    #      at its stated size, plus MSVideo1 inter-frame legs (the headline config is key frames only) ----
    if rank == 0 and world == 1 and args.workload == "c2" and not args.no_codecs:
        codecs = {}
        for name, steps in (("c1", 5), ("c2p", 5), ("c2p8", 5), ("c3", 5), ("c4", 3), ("c5", 2)):
            a2 = argparse.Namespace(**vars(args))
            a2.streams, a2.steps, a2.e2e_steps = 0, min(args.steps, steps), -1
            try:
                l2 = measure(make_workload(name, a2), a2, 0, local_rank, 1, dist, torch, cpu_budget=min(args.cpu_budget, 6.0))
                codecs[name] = {k: l2[k] for k in ("metric", "value", "unit", "steps", "ms_per_step", "config", "kernels", "roofline", "entropy", "e2e", "e2e_inplace", "display_fused", "cpu_baseline") if k in l2}
            except (Exception, SystemExit) as e:       # a failed extra leg must not lose the headline line
                codecs[name] = {"error": str(e)}
        line["codecs"] = codecs

    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
