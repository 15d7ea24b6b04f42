"""Frame side of the per-clip loop: 8-bit decode/normalise and quantise/encode on the device, on-device PSNR/SSIM,
and the 8-bit streamed clip runner.

The reference reads frames with cv2 on the host, normalises them to fp32 there, copies 12 bytes per pixel to the
GPU, copies the fp32 result back and quantises / scores it in numpy (INFN:88-120, 262-276; INF:313-327).  Here only
uint8 crosses PCIe (3 bytes per pixel each way): ``turtle_u8_to_frame`` / ``turtle_frame_to_u8`` convert on the
device, and ``turtle_frame_metrics`` scores the restored frame against its ground truth where it already is.
Everything goes through the C ABI (capi); torch supplies memory and streams.
"""
from __future__ import annotations

import threading
from typing import List, Optional, Sequence

import torch

from . import capi
from .capi import call

_FLAVOURS = {"inference": capi.METRICS_INFERENCE, "basicsr": capi.METRICS_BASICSR, "float": capi.METRICS_FLOAT}


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if not t.is_cuda:
            raise RuntimeError("turtlevsr_b200 runs on CUDA tensors only (no CPU fallback)")


def u8_to_frame(img: torch.Tensor, swap_rb: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """uint8 [H,W,C] (interleaved, as cv2 / PNG decoders deliver it) -> fp32 [C,H,W] in [0,1].  swap_rb: BGR source."""
    _need_cuda(img)
    if img.dtype != torch.uint8 or img.dim() != 3 or img.stride(2) != 1 or img.stride(1) != img.shape[2]:
        raise ValueError("expected a uint8 [H,W,C] image with dense rows")
    H, W, C = img.shape
    if out is None:
        out = torch.empty(C, H, W, device=img.device, dtype=torch.float32)
    with torch.cuda.device(img.device):
        call("turtle_u8_to_frame", img.data_ptr(), img.stride(0), out.data_ptr(), H, W, C, int(swap_rb), _stream(img))
    return out


def frame_to_u8(frame: torch.Tensor, swap_rb: bool = False, round_half_even: bool = True,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 [C,H,W] (or [1,C,H,W]) -> uint8 [H,W,C]: clamp(0,1) * 255, rounded like ``tensor2img`` (img_util.py:99) or
    truncated like ``(x*255).astype(np.uint8)`` (INFN:268-269)."""
    _need_cuda(frame)
    if frame.dim() == 4:
        frame = frame[0]
    frame = frame.float().contiguous()
    C, H, W = frame.shape
    if out is None:
        out = torch.empty(H, W, C, device=frame.device, dtype=torch.uint8)
    with torch.cuda.device(frame.device):
        call("turtle_frame_to_u8", frame.data_ptr(), out.data_ptr(), out.stride(0), H, W, C, int(swap_rb),
             int(round_half_even), _stream(frame))
    return out


class FrameMetrics:
    """PSNR / SSIM of restored frames against ground truth, computed and accumulated on the device.

    ``add`` enqueues two kernels and returns immediately; nothing is copied to the host until ``per_frame`` /
    ``mean`` is called (once per clip), unlike INF:313-327 which converts both frames to numpy every frame.
    flavour: "inference" (INF:33-61), "basicsr" (VRM:171-200 -> metrics/psnr_ssim.py) or "float"."""

    def __init__(self, flavour: str = "inference", capacity: int = 1024):
        self.flavour = _FLAVOURS[flavour]
        self.capacity = capacity
        self.results: Optional[torch.Tensor] = None
        self.ws: Optional[torch.Tensor] = None
        self.n = 0

    def add(self, restored: torch.Tensor, gt: torch.Tensor) -> None:
        _need_cuda(restored, gt)
        if restored.dim() == 4:
            restored = restored[0]
        if gt.dim() == 4:
            gt = gt[0]
        restored, gt = restored.float().contiguous(), gt.float().contiguous()
        if restored.shape != gt.shape:
            raise ValueError(f"Image shapes are differnet: {tuple(restored.shape)}, {tuple(gt.shape)}.")   # psnr_ssim.py:35
        C, H, W = restored.shape
        if self.results is None:
            self.results = torch.zeros(self.capacity, 4, device=restored.device, dtype=torch.float64)
        if self.n >= self.capacity:
            self.results = torch.cat([self.results, torch.zeros_like(self.results)])
            self.capacity *= 2
        need = capi.load().turtle_frame_metrics_workspace(H, W)
        if self.ws is None or self.ws.numel() < need:
            self.ws = torch.empty(need, device=restored.device, dtype=torch.uint8)
        with torch.cuda.device(restored.device):
            call("turtle_frame_metrics", restored.data_ptr(), gt.data_ptr(), C, H, W, self.flavour,
                 self.results[self.n].data_ptr(), self.ws.data_ptr(), _stream(restored))
        self.n += 1

    def per_frame(self) -> torch.Tensor:
        """[n, 4] float64 on the host: PSNR (dB), SSIM, MSE, element count per frame added so far (synchronises)."""
        if self.results is None:
            return torch.zeros(0, 4, dtype=torch.float64)
        return self.results[:self.n].cpu()

    def mean(self):
        r = self.per_frame()
        return float(r[:, 0].mean()), float(r[:, 1].mean())


def frame_metrics(restored: torch.Tensor, gt: torch.Tensor, flavour: str = "inference"):
    """(psnr, ssim) of one frame pair (synchronises; use FrameMetrics inside a loop)."""
    m = FrameMetrics(flavour, capacity=1)
    m.add(restored, gt)
    r = m.per_frame()[0]
    return float(r[0]), float(r[1])


@torch.no_grad()
def run_clip_streamed_u8(net, clip_u8: torch.Tensor, out_u8: Optional[torch.Tensor] = None, device=None,
                         gt_u8: Optional[torch.Tensor] = None, metrics: Optional[FrameMetrics] = None,
                         swap_rb: bool = False, round_half_even: bool = True, k=None, v=None,
                         prev: Optional[torch.Tensor] = None):
    """The cached frame loop (VRM:110-129) over a clip of 8-bit frames in (pinned) host memory.

    clip_u8 [T,H,W,C] uint8 -> out_u8 [T,H',W',C] uint8.  Frame j+1 is uploaded (3 bytes / pixel) on a copy stream
    and normalised on the device while frame j is restored; the restored frame is quantised on the device and its 8-bit
    image travels back on a third stream.  With ``gt_u8`` and ``metrics`` every restored frame is scored on the device
    against its (uploaded) ground truth.  Returns (out_u8, k, v, last_frame)."""
    dev = torch.device(device) if device is not None else next(net.parameters()).device
    T, H, W, C = clip_u8.shape
    up = 4 if getattr(net, "variant", "") == "super" else 1
    Co = getattr(net, "out_channels", C)
    if out_u8 is None:
        out_u8 = torch.empty(T, H * up, W * up, Co, dtype=torch.uint8).pin_memory()
    main = torch.cuda.current_stream(dev)
    h2d, d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    raw = [torch.empty(H, W, C, device=dev, dtype=torch.uint8) for _ in range(3)]
    raw_gt = [torch.empty(H * up, W * up, Co, device=dev, dtype=torch.uint8) for _ in range(3)] if gt_u8 is not None else None
    frames = [torch.empty(1, C, H, W, device=dev) for _ in range(3)]
    q_out = [torch.empty(H * up, W * up, Co, device=dev, dtype=torch.uint8) for _ in range(2)]
    up_done = [torch.cuda.Event() for _ in range(3)]
    slot_free = [torch.cuda.Event() for _ in range(3)]
    q_free = [torch.cuda.Event() for _ in range(2)]

    def upload(j):
        s = j % 3
        with torch.cuda.stream(h2d):
            if j >= 3:
                h2d.wait_event(slot_free[s])
            raw[s].copy_(clip_u8[j], non_blocking=True)
            if raw_gt is not None:
                raw_gt[s].copy_(gt_u8[j], non_blocking=True)
            up_done[s].record(h2d)

    upload(0)
    for j in range(T):
        if j + 1 < T:
            upload(j + 1)
        s = j % 3
        main.wait_event(up_done[s])
        u8_to_frame(raw[s], swap_rb=swap_rb, out=frames[s][0])
        cur = frames[s]
        pre = (cur if prev is None else prev) if j == 0 else frames[(j - 1) % 3]
        o, k, v = net(torch.stack([pre[0], cur[0]], dim=0).unsqueeze(0), k, v)
        if metrics is not None and raw_gt is not None:
            metrics.add(o, u8_to_frame(raw_gt[s], swap_rb=swap_rb))
        if j >= 1:
            slot_free[(j - 1) % 3].record(main)
        qb = j % 2
        if j >= 2:
            main.wait_event(q_free[qb])
        frame_to_u8(o, swap_rb=swap_rb, round_half_even=round_half_even, out=q_out[qb])
        done = torch.cuda.Event()
        done.record(main)
        with torch.cuda.stream(d2h):
            d2h.wait_event(done)
            out_u8[j].copy_(q_out[qb], non_blocking=True)
            q_free[qb].record(d2h)
    d2h.synchronize()
    return out_u8, k, v, frames[(T - 1) % 3].clone()


class FrameFolderReader:
    """Decode the image files of a frame folder (video_to_frames.py's output; INFN:88-120 reads them with cv2.imread
    inside the frame loop) on a background thread into a ring of pinned uint8 buffers, so that PNG decoding overlaps
    the GPU.  Yields uint8 [H,W,C] BGR tensors (cv2 order; pass ``swap_rb=True`` downstream)."""

    def __init__(self, paths: Sequence[str], depth: int = 4):
        import cv2  # host-side decode only; not needed by anything else in the package
        self._cv2 = cv2
        self.paths = list(paths)
        self.depth = depth
        self._slots: List[Optional[torch.Tensor]] = [None] * depth
        self._ready = [threading.Event() for _ in range(depth)]
        self._free = [threading.Event() for _ in range(depth)]
        for e in self._free:
            e.set()
        self._err: Optional[BaseException] = None
        self._t = threading.Thread(target=self._work, daemon=True)
        self._t.start()

    def _work(self):
        try:
            for i, p in enumerate(self.paths):
                s = i % self.depth
                self._free[s].wait()
                self._free[s].clear()
                img = self._cv2.imread(p, self._cv2.IMREAD_COLOR)
                if img is None:
                    raise FileNotFoundError(p)
                t = torch.from_numpy(img)
                if torch.cuda.is_available():
                    if self._slots[s] is None or self._slots[s].shape != t.shape:
                        self._slots[s] = torch.empty_like(t).pin_memory()
                    self._slots[s].copy_(t)
                else:
                    self._slots[s] = t
                self._ready[s].set()
        except BaseException as e:            # surfaced to the consumer
            self._err = e
            for e_ in self._ready:
                e_.set()

    def __len__(self):
        return len(self.paths)

    def __iter__(self):
        for i in range(len(self.paths)):
            s = i % self.depth
            self._ready[s].wait()
            if self._err is not None:
                raise self._err
            self._ready[s].clear()
            yield self._slots[s]
            self._free[s].set()


class FrameFolderWriter:
    """Encode restored 8-bit frames to image files (``cv2.imwrite`` of INFN:272-273) on a background thread, so PNG
    encoding overlaps the GPU.  ``put(path, frame_u8_HWC)`` copies the (pinned) host frame into a small queue and returns;
    ``close()`` drains it.  Frames are expected in cv2's BGR order (``frame_to_u8(..., swap_rb=True)``)."""

    def __init__(self, depth: int = 8):
        import queue
        import cv2
        self._cv2 = cv2
        self._q: "queue.Queue" = queue.Queue(maxsize=depth)
        self._err: Optional[BaseException] = None
        self._t = threading.Thread(target=self._work, daemon=True)
        self._t.start()

    def _work(self):
        while True:
            item = self._q.get()
            if item is None:
                return
            path, img = item
            try:
                if not self._cv2.imwrite(path, img):
                    raise OSError(f"cv2.imwrite failed for {path}")
            except BaseException as e:
                self._err = e

    def put(self, path: str, frame_u8: torch.Tensor) -> None:
        if self._err is not None:
            raise self._err
        self._q.put((path, frame_u8.cpu().numpy().copy()))

    def close(self) -> None:
        self._q.put(None)
        self._t.join()
        if self._err is not None:
            raise self._err
