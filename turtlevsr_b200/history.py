"""HBM-resident history rings for the five history-carrying blocks.

The reference threads its history through the caller as tensors that are ``torch.cat``-ed and
sliced every frame (T1:272-273,286; T1:581-582,610) -- the whole history is re-copied per frame.
Here each history lives in a device buffer with ``M`` frame slots; a new frame is written into
the next slot by the producing kernel itself, and the tensors handed back to the caller are
*views* of the buffer in exactly the reference shapes:

  SAB caches  k [B, Fc, 1, N, 2c]      v [B, Fc, 1, N, ws*ws*c]        (T1:610)
  FHR caches  k [B, head, Fc*c/head, hw]   v likewise                   (T1:286)

The window slides through the buffer; when it reaches the end, the (<=K) live frames are moved to
the front once (amortised cost 1/(M-K) frame copies per frame).  A tensor passed back by the
caller is recognised as "our current window" by its storage pointer, offset, shape and strides;
anything else (``.cpu().to(device)`` round trips of the tiled loop INF:227-237, clones, slices) is
imported by copy into a fresh ring, so the list API stays fully general.

Stale views.  A window view handed out earlier aliases ring memory.  Slots are only ever appended to
between two compactions, so an older view stays intact (and is imported by copy like any foreign
tensor) until the ring compacts; after that its memory has been overwritten.  Every ring therefore
counts its compactions (``epoch``), every view it publishes carries the epoch it was cut in, and a
view of an earlier epoch passed back to the model raises :class:`StaleCacheError` instead of being
read -- callers that keep a history for more than RING_PERIOD frames must ``.clone()`` it (a clone
is an ordinary foreign tensor).  Views that lost their tag (``.detach()``, slicing) are traced back
to their ring through the storage registry and accepted only inside the area written this epoch.
"""
from __future__ import annotations

import os
import weakref
from typing import Optional, Tuple

import torch

RING_SLOTS = int(os.environ.get("TURTLE_RING_SLOTS", "8"))
# Every ring compacts once per RING_PERIOD frames whatever its depth K (slots = K + period by default), so the joint
# state of all rings of a clip -- and with it the set of device addresses a frame touches -- repeats with that period:
# the frame can be replayed from RING_PERIOD CUDA graphs (engine.FrameEngine, cuda_graphs mode).
RING_PERIOD = max(RING_SLOTS - 3, 2)


def default_slots(keep: int) -> int:
    return max(keep + RING_PERIOD, 2 * keep + 2)


_ring_serial = iter(range(1, 1 << 62))
# storage address of every live ring buffer -> ring (finds the ring behind untagged views of its memory)
_by_storage: "weakref.WeakValueDictionary[int, _RingBase]" = weakref.WeakValueDictionary()


class StaleCacheError(RuntimeError):
    """A cache tensor handed back to the model is a view of history-ring memory that has been overwritten since."""


class _RingBase:
    def __init__(self, keep: int, slots: int):
        self.serial = next(_ring_serial)     # never reused (unlike id()): part of the CUDA-graph cache key
        self.keep = keep                     # K = num_frames_tocache of the owning block
        self.slots = max(slots, 2 * keep + 2)
        self.pos = -1                        # newest committed slot
        self.count = 0                       # committed frames in the window (<= keep)
        # signature of the window last handed out (no tensor refs: the views own the ring, not
        # the other way round, so dropping the caches frees the ring without a GC cycle)
        self._sig = None
        self.epoch = 0                       # number of compactions so far (see "Stale views" above)
        # SabRing only: slots [v16_lo, pos] carry a valid fp16 copy of their value rows (None: no slot does)
        self.v16_lo = None

    def _register(self):
        for buf in (self.kbuf, self.vbuf):
            _by_storage[buf.untyped_storage().data_ptr()] = self

    # -- protocol ---------------------------------------------------------------------
    def begin_push(self) -> int:
        """Slot the producing kernels must write the new frame into (compacts first if needed)."""
        if self.pos + 1 >= self.slots:
            self._compact()
            self.epoch += 1
        return self.pos + 1

    def commit(self) -> None:
        self.pos += 1
        self.count = min(self.count + 1, self.keep)

    @property
    def first_live(self) -> int:
        """Oldest slot that takes part in the current frame's attention (before commit)."""
        return self.pos - self.count + 1

    def matches(self, k: torch.Tensor, v: torch.Tensor) -> bool:
        return self._sig is not None and self._sig == (_sig(k), _sig(v))

    def _publish(self, k: torch.Tensor, v: torch.Tensor):
        k._turtle_ring = v._turtle_ring = self
        k._turtle_epoch = v._turtle_epoch = self.epoch
        self._sig = (_sig(k), _sig(v))
        return k, v

    def check_alive(self, t: torch.Tensor) -> None:
        """``t`` aliases this ring's memory but is not its current window: raise unless its bytes are still the
        frames it was cut from."""
        ep = getattr(t, "_turtle_epoch", None)
        if ep is None:
            # untagged view (detach / slice): intact only if it lies inside the slots written since the last compaction
            ok = self._inside_written(t)
        else:
            ok = ep == self.epoch
        if not ok:
            raise StaleCacheError(
                "a history cache passed to the model is a view of ring memory that has been overwritten since it was "
                "returned (the ring compacts every %d frames); .clone() caches that must outlive that" % RING_PERIOD)


def _sig(t: torch.Tensor):
    return (t.device, t.dtype, t.data_ptr(), tuple(t.shape), tuple(t.stride()))


class SabRing(_RingBase):
    """History of a StateAlignBlock: k rows [N,Dk] and v patch rows [N,Dv] per frame."""

    def __init__(self, B, N, Dk, Dv, keep, device, slots=None):
        super().__init__(keep, slots or default_slots(keep))
        self.B, self.N, self.Dk, self.Dv = B, N, Dk, Dv
        self.kbuf = torch.empty(B, self.slots, N, Dk, device=device, dtype=torch.float32)
        self.vbuf = torch.empty(B, self.slots, N, Dv, device=device, dtype=torch.float32)
        # fp16 copy of the value rows for the tensor-core aggregation (csrc/sab_agg_tc.cu): allocated on first use, kept
        # slot for slot next to vbuf; it is engine-internal and never handed to the caller
        self.vbuf16 = None
        self._register()

    def geometry(self):
        return (self.B, self.N, self.Dk, self.Dv)

    def shadow(self) -> torch.Tensor:
        if self.vbuf16 is None:
            self.vbuf16 = torch.empty(self.B, self.slots, self.N, self.Dv, device=self.vbuf.device, dtype=torch.float16)
            self.v16_lo = None
        return self.vbuf16

    def shadow_ok(self) -> bool:
        """Every live slot (the window before this frame's push) has a valid fp16 copy."""
        return self.vbuf16 is not None and (self.count == 0 or (self.v16_lo is not None and self.v16_lo <= self.first_live))

    def _compact(self):
        c, a = self.count, self.first_live
        if c:
            self.kbuf[:, :c].copy_(self.kbuf[:, a:a + c])
            self.vbuf[:, :c].copy_(self.vbuf[:, a:a + c])
            if self.vbuf16 is not None and self.v16_lo is not None:
                self.vbuf16[:, :c].copy_(self.vbuf16[:, a:a + c])
        if self.v16_lo is not None:
            self.v16_lo = max(0, self.v16_lo - a)
        self.pos = c - 1

    def _inside_written(self, t: torch.Tensor) -> bool:
        buf = self.kbuf if t.untyped_storage().data_ptr() == self.kbuf.untyped_storage().data_ptr() else self.vbuf
        per_slot = buf.stride(1)
        off = t.storage_offset() % buf.stride(0)                     # offset inside one batch element
        span = 1 + sum((n - 1) * st for n, st in zip(t.shape[1:], t.stride()[1:])) if t.numel() else 0
        return off + span <= (self.pos + 1) * per_slot

    def views(self):
        a, b = self.first_live, self.pos + 1
        k = self.kbuf[:, a:b].unsqueeze(2)
        v = self.vbuf[:, a:b].unsqueeze(2)
        return self._publish(k, v)

    @classmethod
    def adopt(cls, k: torch.Tensor, v: torch.Tensor, keep: int, device):
        """Import caller-owned cache tensors [B,Fc,1,N,D*] by copy."""
        if k.dim() != 5 or v.dim() != 5 or k.shape[:4] != v.shape[:4] or k.shape[2] != 1:
            raise ValueError(f"StateAlignBlock caches must be [B,F,1,N,D]; got k {tuple(k.shape)}, v {tuple(v.shape)}")
        B, Fc, _, N, Dk = k.shape
        Dv = v.shape[-1]
        r = cls(B, N, Dk, Dv, keep, device)
        Fc = min(Fc, keep)
        r.kbuf[:, :Fc].copy_(k[:, -Fc:, 0].to(device=device, dtype=torch.float32))
        r.vbuf[:, :Fc].copy_(v[:, -Fc:, 0].to(device=device, dtype=torch.float32))
        r.pos, r.count = Fc - 1, Fc
        return r


class FhrRing(_RingBase):
    """History of a FrameHistoryRouter: per frame, normalised key rows and raw value rows, each a
    channels-last map [P, C].  Stored as [B, P, heads, M*ch] so that the reference-shaped window
    [B, heads, Fc*ch, P] is a plain strided view (frames are contiguous inside a head)."""

    def __init__(self, B, P, heads, ch, keep, device, slots=None):
        super().__init__(keep, slots or default_slots(keep))
        self.B, self.P, self.heads, self.ch = B, P, heads, ch
        self.kbuf = torch.empty(B, P, heads, self.slots * ch, device=device, dtype=torch.float32)
        self.vbuf = torch.empty(B, P, heads, self.slots * ch, device=device, dtype=torch.float32)
        self._register()

    def geometry(self):
        return (self.B, self.P, self.heads, self.ch)

    def _inside_written(self, t: torch.Tensor) -> bool:
        # frames are interleaved inside a pixel row ([.., heads, slots*ch]): a view is intact iff its channel range
        # inside a head ends before the first unwritten slot
        off = t.storage_offset() % (self.slots * self.ch)
        width = max((n - 1) * st for n, st in zip(t.shape, t.stride()) if st < self.slots * self.ch) + 1 \
            if t.numel() else 0
        return off + width <= (self.pos + 1) * self.ch

    @property
    def ld(self) -> int:                     # row pitch of one pixel
        return self.heads * self.slots * self.ch

    @property
    def head_stride(self) -> int:
        return self.slots * self.ch

    def _compact(self):
        c, a, ch = self.count, self.first_live, self.ch
        if c:
            self.kbuf[..., :c * ch].copy_(self.kbuf[..., a * ch:(a + c) * ch].clone())
            self.vbuf[..., :c * ch].copy_(self.vbuf[..., a * ch:(a + c) * ch].clone())
        self.pos = c - 1

    def slot_ptr(self, buf: torch.Tensor, b: int, slot: int) -> int:
        from .ops import DevPtr              # an int that remembers its tensor (the custom-op layer needs the owner)
        return DevPtr(buf, 4 * (b * self.P * self.ld + slot * self.ch))

    def views(self):
        a, b, ch = self.first_live, self.pos + 1, self.ch
        k = self.kbuf[..., a * ch:b * ch].permute(0, 2, 3, 1)
        v = self.vbuf[..., a * ch:b * ch].permute(0, 2, 3, 1)
        return self._publish(k, v)

    @classmethod
    def adopt(cls, k: torch.Tensor, v: torch.Tensor, keep: int, ch: int, device):
        if k.dim() != 4 or k.shape != v.shape or k.shape[2] % ch != 0:
            raise ValueError(f"FrameHistoryRouter caches must be [B,heads,F*{ch},hw]; got k {tuple(k.shape)}, "
                             f"v {tuple(v.shape)}")
        B, heads, rows, P = k.shape
        r = cls(B, P, heads, ch, keep, device)
        Fc = min(rows // ch, keep)
        r.kbuf[..., :Fc * ch].copy_(k[:, :, -Fc * ch:].permute(0, 3, 1, 2).to(device=device, dtype=torch.float32))
        r.vbuf[..., :Fc * ch].copy_(v[:, :, -Fc * ch:].permute(0, 3, 1, 2).to(device=device, dtype=torch.float32))
        r.pos, r.count = Fc - 1, Fc
        return r


def _ring_of(t: torch.Tensor):
    ring = getattr(t, "_turtle_ring", None)
    if ring is None:
        ring = _by_storage.get(t.untyped_storage().data_ptr())
    return ring


def resolve_ring(k, v):
    """Ring behind caller-supplied cache tensors if they are exactly its current window (also after ``.detach()``).
    Views of ring memory that are *not* the current window are checked for staleness (StaleCacheError) and then
    treated like any foreign tensor: the caller imports them by copy."""
    if k is None or v is None:
        return None
    rk, rv = _ring_of(k), _ring_of(v)
    if rk is not None and rk is rv and rk.matches(k, v):
        return rk
    if rk is not None:
        rk.check_alive(k)
    if rv is not None:
        rv.check_alive(v)
    return None
