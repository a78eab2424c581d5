"""GPU parity for the mixed corpus path (BASELINE.json configs[4] in miniature): MSVideo1 + ScreenPressor AVI files
-> AVI indexer -> keyframe-delimited segments decoded as independent units -> bit-exact against the oracle decoding
each file as one stream."""
import numpy as np
import pytest

from jsplayer_b200 import avi, BatchDecoder, _lib
import synth
from synth.avi import write_avi
from oracle import pyoracle as O

pytestmark = pytest.mark.gpu


def make_corpus(tmp_path):
    files = []
    w, h = 128, 96
    for i in range(6):
        path = str(tmp_path / ("f%d.avi" % i))
        if i % 3 == 0:
            frames = []
            for f in range(12):
                frames.append(synth.msv1_frame(False, w, h, 100 * i + f, skip_permille=0 if f % 4 == 0 else 400))
            keys = [1 if f % 4 == 0 else 0 for f in range(12)]
            write_avi(path, w, h, 16, b"CRAM", frames, keys)
            files.append((path, O.CODEC_MSVC16, 16, None, frames, keys))
        elif i % 3 == 1:
            pal = synth.random_palette(i)
            frames = [synth.msv1_frame(True, w, h, 100 * i + f, skip_permille=0 if f % 5 == 0 else 300) for f in range(10)]
            keys = [1 if f % 5 == 0 else 0 for f in range(10)]
            write_avi(path, w, h, 8, b"MSVC", frames, keys, palette=pal)
            files.append((path, O.CODEC_MSVC8, 8, pal, frames, keys))
        else:
            frames, keys, _ = synth.sp_stream(w, h, 12, seed=i, version=2 + i % 3, gop=4, change_permille=60)
            write_avi(path, w, h, 24, b"SCPR", frames, keys)
            files.append((path, O.CODEC_SCREENPRESSOR, 24, None, frames, keys))
    return files, w, h


def test_mixed_corpus_gop_sharded(tmp_path):
    files, w, h = make_corpus(tmp_path)
    streams = [avi.load_avi(p, pinned=True) for p, *_ in files]
    specs, where = avi.gop_specs(streams)
    assert len(specs) > len(files)                                    # several segments per file
    bd = BatchDecoder(insignificant_lines=0)
    bd.configure(specs)
    outs, flags = bd.decode_host()
    bd.close()
    assert not (flags & _lib.JSP_FRAME_ERROR).any()
    i = 0
    exp = [O.decode_stream(codec, w, h, bpp, frames, keys=keys, palette=pal)[0] for _, codec, bpp, pal, frames, keys in files]
    for (fi, lo, hi) in where:
        for f in range(lo, hi):
            assert (outs[i] == exp[fi][f]).all(), "file %d frame %d" % (fi, f)
            i += 1
    assert i == sum(len(f[4]) for f in files)


def test_one_shot_multi_gpu_entry_matches(tmp_path):
    """jsp_batch_decode shards whole streams over the visible GPUs (1 here unless the box has more)."""
    import ctypes as C
    files, w, h = make_corpus(tmp_path)
    streams = [avi.load_avi(p) for p, *_ in files]
    specs, where = avi.gop_specs(streams)
    lib = _lib.require_gpu()
    n = len(specs)
    descs = (_lib.StreamDescC * n)()
    keep = []
    total = 0
    for i, sp in enumerate(specs):
        off = np.ascontiguousarray(sp.frame_off, dtype=np.uint64); ln = np.ascontiguousarray(sp.frame_len, dtype=np.uint32)
        keys = np.ascontiguousarray(sp.keys, dtype=np.uint8)
        pal = np.frombuffer(sp.palette, dtype=np.uint8).copy() if sp.palette else None
        keep += [off, ln, keys, pal]
        d = descs[i]
        d.codec, d.width, d.height, d.bpp = int(sp.codec), sp.width, sp.height, sp.bpp
        d.palette = pal.ctypes.data if pal is not None else None
        d.palette_bytes = pal.size if pal is not None else 0
        d.n_frames = len(ln)
        d.bytes = sp.bytes_buf.ctypes.data
        d.frame_off, d.frame_len, d.frame_key = off.ctypes.data, ln.ctypes.data, keys.ctypes.data
        total += len(ln)
    outs = [np.zeros((h, w), dtype=np.int32) for _ in range(total)]
    ptrs = (C.c_void_p * total)(*[o.ctypes.data for o in outs])
    changed = np.zeros(total, dtype=np.uint8); signif = np.zeros(total, dtype=np.uint8); status = np.zeros(total, dtype=np.int32)
    ngpu = max(1, min(2, lib.jsp_device_count()))
    rc = lib.jsp_batch_decode(descs, n, ngpu, ptrs, changed.ctypes.data, signif.ctypes.data, status.ctypes.data)
    assert rc == 0, _lib.last_error()
    assert (status == 0).all()
    exp = [O.decode_stream(codec, w, h, bpp, frames, keys=keys, palette=pal)[0] for _, codec, bpp, pal, frames, keys in files]
    i = 0
    for (fi, lo, hi) in where:
        for f in range(lo, hi):
            assert (outs[i] == exp[fi][f]).all(), "file %d frame %d" % (fi, f)
            i += 1


def test_plain_c_host_decodes_an_avi(tmp_path):
    """examples/decode_avi.c (C99, links only libjsplayer_cuda): AVI file -> GOP segments -> jsp_batch_decode; its
    per-frame checksums must equal the oracle's."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "decode_avi")
    libdir = os.path.join(root, "jsplayer_b200")
    subprocess.run(["gcc", "-std=c99", "-O1", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "decode_avi.c"),
                    "-L", libdir, "-ljsplayer_cuda", "-Wl,-rpath," + libdir, "-o", exe], check=True)
    w, h = 128, 96
    cases = []
    frames, keys, _ = synth.sp_stream(w, h, 9, seed=31, version=4, gop=3, change_permille=60)
    cases.append(("sp.avi", O.CODEC_SCREENPRESSOR, 24, b"SCPR", None, frames, keys))
    pal = synth.random_palette(2)
    frames = [synth.msv1_frame(True, w, h, 50 + f, skip_permille=0 if f % 4 == 0 else 300) for f in range(8)]
    cases.append(("cram8.avi", O.CODEC_MSVC8, 8, b"CRAM", pal, frames, [1 if f % 4 == 0 else 0 for f in range(8)]))
    for name, codec, bpp, four, pal, frames, keys in cases:
        path = str(tmp_path / name)
        write_avi(path, w, h, bpp, four, frames, keys, palette=pal)
        r = subprocess.run([exe, path], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        got = [ln.split() for ln in r.stdout.splitlines() if ln.startswith("frame")]
        exp = O.decode_stream(codec, w, h, bpp, frames, keys=keys, palette=pal)[0]
        assert len(got) == len(frames)
        for f, parts in enumerate(got):
            s = 0
            for v in exp[f].reshape(-1).astype(np.uint32).tolist():
                s = (s * 31 + v) & 0xFFFFFFFF
            assert int(parts[-1], 16) == s, "%s frame %d" % (name, f)
            assert int(parts[parts.index("status") + 1]) == 0


def test_gop_split_with_flat_key_frames_matches_in_order_decoding(tmp_path):
    """A ScreenPressor file whose GOPs begin with FLAT key frames: in stream order the reference decodes them with the
    entropy coder an earlier coded key frame created (ScreenPressor.hx:132-155, :160-162).  Cut into segments, the
    segment has to be told (jsp_segment_stream -> jsp_stream_desc.sp_version), or the flat frame is an error and every
    P frame after it degrades to copy-previous."""
    w, h = 96, 64
    for version in (2, 3, 4):
        enc = synth.SPEncoder(w, h, 24, version)
        p0 = synth.screen(w, h, 1)
        frames, keys = [enc.iframe(p0)], [1]
        cur = p0
        for g in range(3):
            nxt, mv = synth.screen_next(cur, 10 + g, 80)
            frames.append(enc.pframe(nxt, cur, mv)); keys.append(0)
            flat = np.full((h, w), 0x102030 * (g + 1), dtype=np.int32)
            frames.append(enc.flat(0x102030 * (g + 1))); keys.append(1)            # a GOP that starts with a flat key frame
            cur, mv = synth.screen_next(flat, 20 + g, 120)
            frames.append(enc.pframe(cur, flat, mv)); keys.append(0)
            nxt, mv = synth.screen_next(cur, 30 + g, 120)
            frames.append(enc.pframe(nxt, cur, mv)); keys.append(0)
            cur = nxt
        path = str(tmp_path / ("flat%d.avi" % version))
        write_avi(path, w, h, 24, b"SCPR", frames, keys)
        st = avi.load_avi(path, pinned=True)
        segs = st.segments()
        assert len(segs) == 4 and [v for _, _, v in segs] == [0, version, version, version]
        specs, where = avi.gop_specs([st])
        bd = BatchDecoder()
        bd.configure(specs)
        outs, flags = bd.decode_host()
        bd.close()
        exp, ch, sg, stt = O.decode_stream(O.CODEC_SCREENPRESSOR, w, h, 24, frames, keys=keys)
        assert (stt == 0).all() and not (flags & _lib.JSP_FRAME_ERROR).any()
        for i in range(len(frames)):
            assert (outs[i] == exp[i]).all(), "version %d frame %d" % (version, i)


def _one_shot(specs, n_gpus, shapes):
    import ctypes as C
    lib = _lib.require_gpu()
    n = len(specs)
    descs = (_lib.StreamDescC * n)()
    keep, total = [], 0
    for i, sp in enumerate(specs):
        off = np.ascontiguousarray(sp.frame_off, dtype=np.uint64); ln = np.ascontiguousarray(sp.frame_len, dtype=np.uint32)
        keys = np.ascontiguousarray(sp.keys, dtype=np.uint8)
        pal = np.frombuffer(sp.palette, dtype=np.uint8).copy() if sp.palette else None
        keep += [off, ln, keys, pal]
        d = descs[i]
        d.codec, d.width, d.height, d.bpp = int(sp.codec), sp.width, sp.height, sp.bpp
        d.palette = pal.ctypes.data if pal is not None else None
        d.palette_bytes = pal.size if pal is not None else 0
        d.n_frames = len(ln)
        d.bytes = sp.bytes_buf.ctypes.data
        d.frame_off, d.frame_len, d.frame_key = off.ctypes.data, ln.ctypes.data, keys.ctypes.data
        d.sp_version = int(sp.sp_version)
        total += len(ln)
    outs = [np.zeros(shapes[i], dtype=np.int32) for i in range(total)]
    ptrs = (C.c_void_p * total)(*[o.ctypes.data for o in outs])
    changed = np.zeros(total, dtype=np.uint8); signif = np.zeros(total, dtype=np.uint8); status = np.zeros(total, dtype=np.int32)
    rc = lib.jsp_batch_decode(descs, n, n_gpus, ptrs, changed.ctypes.data, signif.ctypes.data, status.ctypes.data)
    assert rc == 0, _lib.last_error()
    return outs, changed, signif, status


def test_multi_gpu_outputs_equal_the_single_gpu_outputs(tmp_path):
    """"2/4/8-GPU outputs must equal the 1-GPU output byte-for-byte".  The mini corpus plus a 1080p pair of
    files (MSVideo1 + ScreenPressor, key frame every 4), cut into GOP segments and decoded by jsp_batch_decode on 1 GPU and
    on every power-of-two GPU count the box has.  Pictures, changed / significant flags and status must be identical (and
    the 1-GPU pictures equal the oracle's).  Skipped on a single-GPU box (run it with `gpurun --gpus 2`)."""
    lib = _lib.require_gpu()
    ndev = lib.jsp_device_count()
    if ndev < 2:
        pytest.skip("needs at least 2 GPUs")
    files, w, h = make_corpus(tmp_path)
    W, H = 1920, 1080
    big = []
    frames = [synth.msv1_frame(False, W, H, 900 + f, skip_permille=0 if f % 4 == 0 else 700) for f in range(12)]
    keys = [1 if f % 4 == 0 else 0 for f in range(12)]
    p = str(tmp_path / "big_m.avi"); write_avi(p, W, H, 16, b"CRAM", frames, keys)
    big.append((p, O.CODEC_MSVC16, 16, None, frames, keys))
    frames, keys, _ = synth.sp_stream(W, H, 12, seed=77, version=4, gop=4, change_permille=30)
    p = str(tmp_path / "big_s.avi"); write_avi(p, W, H, 24, b"SCPR", frames, keys)
    big.append((p, O.CODEC_SCREENPRESSOR, 24, None, frames, keys))
    allfiles = files + big
    streams = [avi.load_avi(p, pinned=True) for p, *_ in allfiles]
    specs, where = avi.gop_specs(streams)
    shapes = []
    for (fi, lo, hi) in where:
        shapes += [(streams[fi].height, streams[fi].width)] * (hi - lo)
    ref = _one_shot(specs, 1, shapes)
    exp = [O.decode_stream(codec, streams[k].width, streams[k].height, bpp, frames, keys=keys, palette=pal)[0]
           for k, (_, codec, bpp, pal, frames, keys) in enumerate(allfiles)]
    i = 0
    for (fi, lo, hi) in where:
        for f in range(lo, hi):
            assert (ref[0][i] == exp[fi][f]).all(), "1 GPU: file %d frame %d" % (fi, f)
            i += 1
    g = 2
    while g <= ndev:
        got = _one_shot(specs, g, shapes)
        for i in range(len(shapes)):
            assert got[0][i].tobytes() == ref[0][i].tobytes(), "%d GPUs: picture %d differs from the 1-GPU picture" % (g, i)
        for k in (1, 2, 3):
            assert (got[k] == ref[k]).all()
        g *= 2
