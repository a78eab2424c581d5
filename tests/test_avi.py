"""AVI indexer (csrc/avi_index.cpp behind jsp_avi_*) and the GOP / shard logic of jsplayer_b200/avi.py: host-side
tests, no GPU.  Format facts follow reference src/AVIParser.hx:42-184, src/DataLoaderAVIIndexed.hx:276-350 and
src/DataLoader.hx:321-401."""
import struct

import numpy as np
import pytest

from jsplayer_b200 import avi, CodecType
import synth
from synth.avi import avi_bytes, _chunk, _list


def parse(b):
    return avi.parse_avi(np.frombuffer(b, dtype=np.uint8))


def sp_frames(n=6, gop=3, version=2):
    return synth.sp_stream(64, 48, n, seed=1, version=version, gop=gop)[:2]


def test_idx1_relative_offsets_and_payload_lengths():
    fr, k = sp_frames()
    a = parse(avi_bytes(64, 48, 24, b"SCPR", fr, k, fps=25))
    assert (a.codec, a.width, a.height, a.bpp, a.n_frames, a.has_index) == (CodecType.codec_screenpressor, 64, 48, 24, 6, True)
    assert abs(a.fps - 25) < 1e-3
    assert list(a.keys) == k
    assert all(bytes(a.frame(i)) == fr[i] for i in range(6))           # odd-sized frames come without the RIFF pad byte
    assert a.gops() == [(0, 3), (3, 6)]


def test_idx1_absolute_offsets():
    fr, k = sp_frames()
    b = bytearray(avi_bytes(64, 48, 24, b"SCPR", fr, k))
    movi = b.find(b"movi")
    idx = b.find(b"idx1")
    n = struct.unpack_from("<I", b, idx + 4)[0] // 16
    for i in range(n):                                                  # rewrite the offsets as absolute file positions
        o = idx + 8 + 16 * i + 8
        struct.pack_into("<I", b, o, struct.unpack_from("<I", b, o)[0] + movi)
    a = parse(bytes(b))
    assert list(a.keys) == k and all(bytes(a.frame(i)) == fr[i] for i in range(6))


def test_codec_choice_and_palette():
    pal = synth.random_palette(3)
    fr = [synth.msv1_frame(True, 64, 48, 1)] + [synth.msv1_frame(True, 64, 48, 2 + i, skip_permille=300) for i in range(3)]
    for four in (b"CRAM", b"MSVC", b"msvc"):
        a = parse(avi_bytes(64, 48, 8, four, fr, None, palette=pal))
        assert a.codec == CodecType.codec_msvc8 and a.palette == pal and list(a.keys) == [1, 0, 0, 0]
    fr16 = [synth.msv1_frame(False, 64, 48, 1)]
    assert parse(avi_bytes(64, 48, 16, b"CRAM", fr16)).codec == CodecType.codec_msvc16
    assert parse(avi_bytes(64, 48, 24, b"SCPR", *sp_frames())).codec == CodecType.codec_screenpressor


def strip_idx1(b):
    i = b.find(b"idx1")
    body = b[8:i]
    return b"RIFF" + struct.pack("<I", len(body)) + body


def test_no_index_falls_back_to_the_codecs_is_key_frame():
    fr, k = sp_frames(version=4)
    a = parse(strip_idx1(avi_bytes(64, 48, 24, b"SCPR", fr, k)))
    assert not a.has_index and list(a.keys) == k                       # ScreenPressor.hx:96-101
    frm = [synth.msv1_frame(False, 64, 48, 1)] + [synth.msv1_frame(False, 64, 48, 2 + i, skip_permille=300) for i in range(3)]
    a = parse(strip_idx1(avi_bytes(64, 48, 16, b"CRAM", frm)))
    assert list(a.keys) == [1, 0, 0, 0]                                  # MSVideo1.hx:226-259: key <=> no skip opcode


def test_opendml_ix00_and_list_rec():
    """ix00 standard index inside movi (key = bit 31 of the size clear, DataLoader.hx:338-347) and frames wrapped in
    LIST 'rec ' (AVIParser.hx:150)."""
    fr, k = sp_frames()
    whole = avi_bytes(64, 48, 24, b"SCPR", fr, k)
    head = whole[12:whole.find(b"movi") - 8]                             # everything before LIST movi
    movi_payload = b""
    positions = []
    pos0 = 12 + len(head) + 12                                           # file offset of the first byte after 'movi'
    for i, f in enumerate(fr):
        c = _chunk(b"00dc", bytes(f))
        if i % 2 == 0:
            positions.append(pos0 + len(movi_payload) + 12 + 8)          # LIST size 'rec ' + chunk header
            c = _list(b"rec ", c)
        else:
            positions.append(pos0 + len(movi_payload) + 8)
        movi_payload += c
    base = pos0
    ents = b"".join(struct.pack("<II", p - base, len(f) | (0 if kk else 0x80000000)) for p, f, kk in zip(positions, fr, k))
    ix = struct.pack("<HBBI4sQI", 2, 0, 1, len(fr), b"00dc", base, 0) + ents
    movi_payload += _chunk(b"ix00", ix)
    body = b"AVI " + head + _list(b"movi", movi_payload)
    a = parse(b"RIFF" + struct.pack("<I", len(body)) + body)
    assert a.has_index and a.n_frames == 6
    assert list(a.keys) == k and all(bytes(a.frame(i)) == fr[i] for i in range(6))


def test_bad_files():
    with pytest.raises(ValueError):
        parse(b"not an avi file at all")
    with pytest.raises(ValueError):
        parse(b"RIFF\x04\0\0\0AVI ")
    fr, k = sp_frames()
    b = avi_bytes(64, 48, 24, b"SCPR", fr, k)
    a = parse(b[: len(b) // 2])                                           # truncated: the frames that are there
    assert 0 < a.n_frames <= 6


def test_mutated_files_never_leave_the_buffer():
    """Bit flips, random / extreme 32-bit fields and truncation anywhere in the file (headers, chunk sizes, both index
    kinds): the indexer either refuses the file or returns a frame table that lies inside it."""
    rng = np.random.default_rng(5)
    fr = [synth.msv1_frame(True, 64, 48, s, skip_permille=100 if s else 0) for s in range(6)]
    seeds = [avi_bytes(64, 48, 8, b"MSVC", fr, keys=[1, 0, 0, 1, 0, 0], palette=synth.random_palette(3))]
    fr2, k2 = sp_frames()
    plain = avi_bytes(64, 48, 24, b"SCPR", fr2, k2)
    movi = plain.find(b"movi")
    pos, ents = movi + 4, []
    while plain[pos:pos + 4] == b"00dc":
        ln = struct.unpack("<I", plain[pos + 4:pos + 8])[0]
        ents.append((pos + 8, ln)); pos += 8 + ((ln + 1) & ~1)
    ix = struct.pack("<HBBI4sQI", 2, 0, 1, len(ents), b"00dc", 0, 0) + b"".join(
        struct.pack("<II", o, l | (0x80000000 if i % 3 else 0)) for i, (o, l) in enumerate(ents))
    seeds.append(plain[:pos] + _chunk(b"ix00", ix) + plain[pos:])          # stale LIST size: also a malformed file
    for base in seeds:
        base = np.frombuffer(base, dtype=np.uint8)
        accepted = 0
        for it in range(400):
            m = base.copy()
            for _ in range(1 + it % 7):
                at = int(rng.integers(0, m.size))
                if it % 4 == 0:
                    m[at] ^= 1 << int(rng.integers(0, 8))
                elif it % 4 == 1:
                    m[at:at + 4] = rng.integers(0, 256, min(4, m.size - at), dtype=np.uint8)
                elif it % 4 == 2:
                    v = 0xFFFFFFFF if rng.random() < 0.3 else int(rng.integers(0, 64))
                    a4 = at & ~3
                    if a4 + 4 <= m.size:
                        m[a4:a4 + 4] = np.frombuffer(np.uint32(v).tobytes(), dtype=np.uint8)
                else:
                    m = m[: 12 + int(rng.integers(0, m.size - 12))]
                    break
            try:
                a = avi.parse_avi(m)
            except ValueError:
                continue
            accepted += 1
            assert all(int(a.frame_off[i]) + int(a.frame_len[i]) <= m.size for i in range(a.n_frames))
            a.gops(); a.spec()
        assert accepted > 100


def test_shard_is_a_balanced_partition():
    w = [5, 3, 8, 1, 9, 2, 7, 7]
    for n in (1, 2, 3, 8):
        parts = avi.shard(w, n)
        assert sorted(i for p in parts for i in p) == list(range(len(w)))
        loads = [sum(w[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(w)


def _rank_main(rank, world, port, q):
    """One rank of the stream-sharded job on CPU: same partition on every rank, own shard only, no data exchange --
    the only collectives are the ones bench.py uses (barrier + MAX of the elapsed time)."""
    import os
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    streams = []
    for s in range(5):
        fr, k = synth.sp_stream(48, 32, 6, seed=10 + s, version=2 + s % 3, gop=2 + s % 2)[:2]
        streams.append(avi.parse_avi(np.frombuffer(avi_bytes(48, 32, 24, b"SCPR", fr, k), dtype=np.uint8)))
    specs, where = avi.gop_specs(streams)
    weights = [int(sp.frame_len.sum()) + sp.width * sp.height * sp.n_frames // 8 for sp in specs]
    mine = avi.shard(weights, world)[rank]
    dist.barrier()
    t = torch.tensor([float(len(mine))], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    got = [None] * world
    dist.all_gather_object(got, mine)
    q.put((rank, mine, got, len(specs), float(t.item())))
    dist.destroy_process_group()


def test_two_ranks_shard_gops_without_exchange():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (np.random.default_rng().integers(0, 2000))
    ps = [ctx.Process(target=_rank_main, args=(r, 2, int(port), q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    n = res[0][3]
    for rank, mine, got, n_specs, tmax in res:
        assert got[rank] == mine and n_specs == n
        assert sorted(got[0] + got[1]) == list(range(n))                 # every GOP decoded exactly once
        assert tmax == max(len(got[0]), len(got[1]))


def test_segmenter_carries_the_screenpressor_coder_version():
    """jsp_segment_stream: segments start at key frames; ScreenPressor's entropy coder is created by the first coded key
    frame that names a known version and kept (ScreenPressor.hx:160-162), so later segments inherit it."""
    import ctypes as C
    from jsplayer_b200 import _lib
    lib = _lib.load()
    frames = [b"\x99", b"\x11abc", b"\x01zz", b"\x52", b"\x22abcdef", b"\x01", b"\x11abc", b"\x01", b"\x32xyz", b""]
    keys = [0, 1, 0, 1, 1, 0, 1, 0, 1, 1]
    blob = np.frombuffer(b"".join(frames) + b"\0", dtype=np.uint8).copy()
    ln = np.array([len(f) for f in frames], dtype=np.uint32)
    off = np.concatenate([[0], np.cumsum(ln)[:-1]]).astype(np.uint64)
    k = np.array(keys, dtype=np.uint8)
    first = np.zeros(len(frames), dtype=np.int32); ver = np.zeros(len(frames), dtype=np.int32)
    n = lib.jsp_segment_stream(int(CodecType.codec_screenpressor), blob.ctypes.data, off.ctypes.data, ln.ctypes.data,
                               k.ctypes.data, len(frames), first.ctypes.data, ver.ctypes.data)
    assert n == 7
    assert list(first[:n]) == [0, 1, 3, 4, 6, 8, 9]
    # flat 0x11 and the unknown-version 0x52 create no coder; 0x22 (version 3) does, and 0x32 later does not replace it
    assert list(ver[:n]) == [0, 0, 0, 0, 3, 3, 3]
    n = lib.jsp_segment_stream(int(CodecType.codec_msvc16), blob.ctypes.data, off.ctypes.data, ln.ctypes.data,
                               k.ctypes.data, len(frames), first.ctypes.data, ver.ctypes.data)
    assert n == 7 and not ver[:n].any()
