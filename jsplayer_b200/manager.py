"""Headless mirror of the reference's frame scheduler around the codec: `Manager.worker` and its buffer ring.

The reference drives an IVideoCodec from `Manager` (src/Manager.hx): `num_buffers` Int32Array pictures whose ownership is
tracked as `trash | has_frames(first, last)` (:114-118, :568-578), `get_free_buffer` that never hands out the buffer the codec
still borrows as its previous frame (:424-443), `worker` that decodes the next frame into a free buffer and books the result
(:454-539), a restart from the nearest key frame with every buffer trashed when the frame of interest lies outside the decoded
run (:244-249), and `SkipStills` over the per-frame `significant_changes` (:289-317, DataLoader.hx:239-252).  This module is
that protocol, member names kept, with everything that is not the codec path removed (timers, the loaders' HTTP state machine,
audio, the canvas): frames are all "loaded", time is a frame index.  It takes ANY decoder with the IVideoCodec members -- the
GPU drop-ins of jsplayer_b200.codec or a CPU implementation -- so the same script of plays, seeks and still-skips can be run
against both and compared buffer by buffer (tests/test_manager_gpu.py).  (the literal Haxe driver needs a Haxe
toolchain; this is its language-neutral restatement.)
"""
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from .codec import DecoderState

INSIGNIFICANT_LINES = 36                       # Manager.hx:61
TRASH = None                                   # BufferState.trash (Manager.hx:114-118); has_frames = (first, last)


@dataclass
class CompressedFrame:                         # VideoData.hx:68-73
    key: bool
    data: bytes
    significant_changes: Optional[bool] = None


class Manager:
    def __init__(self, decoder, width, height, frames: Sequence[bytes], keys: Sequence[int], nbuffers=9, fps=1.0):
        """`decoder`: an IVideoCodec (Preinit is called here, Manager.hx:128).  nbuffers: Main creates Manager(9)."""
        self.decoder = decoder
        self.X, self.Y = int(width), int(height)
        self.num_buffers = int(nbuffers)
        self.buffers = [np.zeros(self.X * self.Y, dtype=np.int32) for _ in range(self.num_buffers + 1)]    # + one for conversion (:117-119)
        self.bufs: List[Optional[tuple]] = [TRASH] * self.num_buffers
        self.frames = [CompressedFrame(bool(k), bytes(f)) for f, k in zip(frames, keys)]
        self.nframes = len(self.frames)
        self.fps = float(fps)
        self.next_frame_to_decode = 0
        self.frame_of_interest = 0
        self.decoded_log = []                  # (buffer index, frame number) in the order `decoded` fired (:577)
        decoder.Preinit(INSIGNIFICANT_LINES)

    # ---- DataLoader pieces the scheduler needs (everything is loaded) ----
    def GetNearestKeyframe(self, n):           # DataLoader.hx:125-132
        if not self.frames:
            return 0
        n = min(n, len(self.frames) - 1)
        while not self.frames[n].key and n > 0:
            n -= 1
        return n

    def FindPossibleChange(self, pos_from):    # DataLoader.hx:239-252 -> ("change" | "unknown", index)
        for i in range(pos_from, len(self.frames)):
            ch = self.frames[i].significant_changes
            if ch is None:
                return ("unknown", i)
            if ch:
                return ("change", i)
        return ("change", len(self.frames) - 1) if self.frames else ("unknown", pos_from)

    # ---- Manager ----
    def get_free_buffer(self, prev_frame_buf_index):          # Manager.hx:424-443; -1 = no buffer available
        oldest_index, oldest_frame = -1, 100000000
        for i in range(len(self.bufs)):
            if i == prev_frame_buf_index:
                continue
            st = self.bufs[i]
            if st is TRASH:
                return i
            first, last = st
            if last < self.frame_of_interest and first < oldest_frame:
                oldest_frame, oldest_index = first, i
        if oldest_index >= 0:
            self.bufs[oldest_index] = TRASH
            return oldest_index
        return -1

    def update_bufs(self, idx, frame_num, new_data):          # Manager.hx:568-578
        st = self.bufs[idx]
        if st is TRASH or new_data or st[1] != frame_num - 1:
            self.bufs[idx] = (frame_num, frame_num)
        else:
            self.bufs[idx] = (st[0], frame_num)
        self.decoded_log.append((idx, frame_num))

    def frames_differ_significantly(self, new_frame, prev_frame, cur):   # Manager.hx:392-421
        n = self.next_frame_to_decode
        if n > 0:
            frm = self.frames[n - 1]
            if frm.key and frm.data is not None:
                return frm.data != cur.data                   # two key frames: byte compare, different lengths differ
        else:
            return True
        if prev_frame is None:                                # `pnt2[i]` of null would throw in the reference; cannot happen after frame 0
            return True
        a = INSIGNIFICANT_LINES * self.X
        return bool((new_frame[a:self.X * self.Y] != prev_frame[a:self.X * self.Y]).any())

    def _index_of(self, arr):
        if arr is None:
            return -1
        for i, b in enumerate(self.buffers):
            if b is arr:
                return i
        return -1

    def worker(self):                                         # Manager.hx:454-539 (no timers, nothing loading)
        """One step: decode frame `next_frame_to_decode` if a buffer is free.  Returns False when nothing could be done."""
        if self.next_frame_to_decode >= self.nframes:
            return False
        prev_frame = self.decoder.PreviousFrame()
        prev_idx = self._index_of(prev_frame)
        free_idx = self.get_free_buffer(prev_idx)
        if free_idx < 0:
            return False                                      # no free bufs to decode to (:474-478)
        frm = self.frames[self.next_frame_to_decode]
        new_frame = self.buffers[free_idx]
        if frm.key:
            state = self.decoder.DecompressI(frm.data, new_frame)
            if state == DecoderState.zero_state:              # handle_decode_status -> on_idecoded (:445-452, :499-504)
                self.update_bufs(free_idx, self.next_frame_to_decode, True)
                if frm.significant_changes is None:
                    frm.significant_changes = self.frames_differ_significantly(new_frame, prev_frame, frm)
                self.next_frame_to_decode += 1
            else:
                # error_occured: the reference only traces and retries the same frame for ever; a headless run has to move on
                self.next_frame_to_decode += 1
        else:
            data_pnt, significant = self.decoder.DecompressP(frm.data, new_frame)      # PFrameResult{data_pnt, significant_changes}
            frm.significant_changes = bool(significant)
            if data_pnt is not None:                          # do nothing if no meaningful data decoded (:515)
                if data_pnt is prev_frame:
                    self.update_bufs(prev_idx, self.next_frame_to_decode, False)
                else:
                    self.update_bufs(free_idx, self.next_frame_to_decode, True)
            self.next_frame_to_decode += 1
        return True

    def GetDecompressedFrame(self, frame):                    # Manager.hx:214-259 with time = frame index
        """The buffer holding `frame`, or None ("soon": the worker has to run; a seek restarts it at the key frame)."""
        self.frame_of_interest = int(frame)
        for nb, st in enumerate(self.bufs):
            if st is not TRASH and st[0] <= self.frame_of_interest <= st[1]:
                return self.buffers[nb]
        key_idx = self.GetNearestKeyframe(self.frame_of_interest)
        if self.next_frame_to_decode < key_idx or self.next_frame_to_decode > self.frame_of_interest:      # seek (:244-249)
            self.next_frame_to_decode = key_idx
            for i in range(len(self.bufs)):
                self.bufs[i] = TRASH
        return None

    def show(self, frame, max_steps=100000):
        """GetDecompressedFrame + as many worker steps as it takes (what the timer does)."""
        for _ in range(max_steps):
            b = self.GetDecompressedFrame(frame)
            if b is not None:
                return b
            if not self.worker():
                return None
        return None

    def SkipStills(self, first_call=True, max_steps=100000):  # Manager.hx:289-317 without the think-time limit
        """Advances frame_of_interest to the next frame with significant changes; returns its index."""
        if first_call:
            self.frame_of_interest += 1
        for _ in range(max_steps):
            kind, pos = self.FindPossibleChange(self.frame_of_interest)
            self.frame_of_interest = pos
            if kind == "change":
                return pos
            while self.next_frame_to_decode <= self.frame_of_interest:
                if not self.worker():
                    # every buffer holds frames at or after the frame of interest: the player would draw and move on
                    self.GetDecompressedFrame(self.frame_of_interest)
                    if not self.worker():
                        return None
        return None
