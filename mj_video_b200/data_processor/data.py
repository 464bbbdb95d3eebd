"""Drop-in for the frame path of scripts/data_processor/data.py (load_video :158-179, get_index :127-137,
dynamic_preprocess :81-117, build_transform :56-64).

`load_video` keeps the reference signature and return value (CPU fp32 `pixel_values [sum(tiles),3,S,S]`,
`num_patches_list`). The per-frame work — PIL bicubic resize to the tile grid, row-major crops, thumbnail, ToTensor,
Normalize — runs in one CUDA pass over the selected uint8 frames (mjv_preprocess_frames), bit-exact with Pillow +
torchvision. Two keyword-only extensions keep the tiles on the GPU for the scoring path: `device=` and `dtype=`.

Decoding is delegated: `decord` (as the reference) or PyAV when installed, `.npy`/`.npz` uint8 [F,H,W,3] arrays, or an
in-memory array/tensor (SURVEY §8(f) N1 lists NVDEC decode as the next component, not part of this path).
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np
import torch

from .. import _lib

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def get_index(bound, fps, max_frame, first_idx=0, num_segments=32):
    """data.py:127-137 — uniformly spaced frame indices (float64 linspace truncated toward zero)."""
    if bound is None:
        return np.linspace(first_idx, max_frame, num_segments, endpoint=False, dtype=int)
    start_time, end_time = bound
    return np.linspace(int(start_time * fps), int(end_time * fps), num_segments, endpoint=False, dtype=int)


def find_closest_aspect_ratio(width, height, min_num=1, max_num=12, image_size=448):
    """Tile grid (cols, rows) chosen by dynamic_preprocess (data.py:66-99); host-side helper of the CUDA library."""
    gi, gj = C.c_int32(), C.c_int32()
    _lib.check(_lib.load().mjv_choose_grid(int(width), int(height), int(min_num), int(max_num), int(image_size),
                                           C.byref(gi), C.byref(gj)))
    return gi.value, gj.value


class _ArrayReader:
    def __init__(self, frames, fps=8.0):
        self.frames = frames
        self.fps = fps

    def __len__(self):
        return len(self.frames)

    def get_avg_fps(self):
        return self.fps

    def take(self, indices) -> np.ndarray:
        f = self.frames
        if isinstance(f, torch.Tensor):
            return f[torch.as_tensor(np.asarray(indices))].cpu().numpy()
        return np.stack([np.asarray(f[int(i)]) for i in indices])


class _DecordReader:
    def __init__(self, path):
        from decord import VideoReader, cpu
        self.vr = VideoReader(str(path), ctx=cpu(0), num_threads=1)

    def __len__(self):
        return len(self.vr)

    def get_avg_fps(self):
        return float(self.vr.get_avg_fps())

    def take(self, indices) -> np.ndarray:
        return np.stack([self.vr[int(i)].asnumpy() for i in indices])


class _AvReader:
    def __init__(self, path):
        import av
        self.container = av.open(str(path))
        stream = self.container.streams.video[0]
        self.frames = [f.to_ndarray(format="rgb24") for f in self.container.decode(stream)]
        self.fps = float(stream.average_rate) if stream.average_rate else 25.0

    def __len__(self):
        return len(self.frames)

    def get_avg_fps(self):
        return self.fps

    def take(self, indices) -> np.ndarray:
        return np.stack([self.frames[int(i)] for i in indices])


class _OpenCVReader:
    """cv2.VideoCapture fallback (the container ships OpenCV but neither decord nor PyAV). Frames are decoded in index
    order while walking the stream once — robust for codecs whose random seeking is inexact."""

    def __init__(self, path):
        import cv2
        self.cv2 = cv2
        self.path = str(path)
        cap = cv2.VideoCapture(self.path)
        if not cap.isOpened():
            raise RuntimeError(f"OpenCV cannot open {self.path}")
        self.n = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
        self.fps = float(cap.get(cv2.CAP_PROP_FPS)) or 25.0
        cap.release()
        if self.n <= 0:
            raise RuntimeError(f"{self.path}: no frames")

    def __len__(self):
        return self.n

    def get_avg_fps(self):
        return self.fps

    def take(self, indices) -> np.ndarray:
        cv2 = self.cv2
        want = [int(i) for i in indices]
        order = sorted(set(want))
        got = {}
        cap = cv2.VideoCapture(self.path)
        pos = 0
        for target in order:
            while pos < target:       # skip without decoding to RGB
                if not cap.grab():
                    break
                pos += 1
            ok, frame = cap.read()
            pos += 1
            if not ok:
                raise RuntimeError(f"{self.path}: cannot decode frame {target}")
            got[target] = cv2.cvtColor(frame, cv2.COLOR_BGR2RGB)
        cap.release()
        return np.stack([got[i] for i in want])


def open_video(video):
    """Frame source with `len`, `get_avg_fps`, `take(indices) -> uint8 [n,H,W,3]`."""
    if isinstance(video, (np.ndarray, torch.Tensor)):
        return _ArrayReader(video)
    path = Path(str(video))
    if str(video).startswith("http"):
        raise RuntimeError("URL download (data.py:33-54) is outside this path; pass a local file")
    if path.suffix == ".npy":
        return _ArrayReader(np.load(path, mmap_mode="r"))
    if path.suffix == ".npz":
        z = np.load(path)
        return _ArrayReader(z[z.files[0]])
    try:
        return _DecordReader(path)
    except ImportError:
        pass
    try:
        return _AvReader(path)
    except ImportError:
        pass
    try:
        return _OpenCVReader(path)
    except ImportError as e:
        raise RuntimeError("no video decoder available (decord / av / cv2 not installed); pass an array or .npy") from e


def sample_clip(video, bound=None, num_segments=32) -> np.ndarray:
    """The host half of load_video (data.py:164-172): open, `get_index`, decode ONLY the sampled frames -> uint8 [n,H,W,3]."""
    vr = open_video(video)
    idx = get_index(bound, float(vr.get_avg_fps()), len(vr) - 1, first_idx=0, num_segments=num_segments)
    frames = np.ascontiguousarray(vr.take(idx))
    if frames.ndim != 4 or frames.shape[-1] != 3:
        raise ValueError(f"expected RGB frames [n,H,W,3], got {frames.shape}")
    return frames


class ClipDecodePool:
    """Host decode of many clips in parallel (the part that does not need NVDEC). The reference decodes
    one video at a time on one thread (`VideoReader(..., num_threads=1)`, data.py:164) - about 30 clips/s at best, while one
    B200 consumes 85 clips/s and a box of eight ~650. decord, PyAV and OpenCV all release the GIL while decoding, so a
    thread pool scales with the host cores; results come back in submission order and are bit-identical to
    `sample_clip` called serially (same readers, same `get_index`)."""

    def __init__(self, workers: int = 8, num_segments: int = 8, bound=None):
        from concurrent.futures import ThreadPoolExecutor
        self.num_segments, self.bound = num_segments, bound
        self._ex = ThreadPoolExecutor(max_workers=max(1, int(workers)), thread_name_prefix="mjv-decode")

    def submit(self, video):
        return self._ex.submit(sample_clip, video, self.bound, self.num_segments)

    def map(self, videos):
        """Decode `videos` concurrently; yields uint8 clips in input order (a failed clip raises when it is reached)."""
        futs = [self.submit(v) for v in videos]
        for f in futs:
            yield f.result()

    def close(self):
        self._ex.shutdown(wait=True, cancel_futures=True)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def preprocess_frames(frames_u8: torch.Tensor, input_size=448, max_num=1, use_thumbnail=True,
                      dtype=torch.float32, out: torch.Tensor | None = None) -> tuple[torch.Tensor, int]:
    """uint8 CUDA tensor [F,H,W,3] -> normalised tiles [F*tiles_per_frame,3,S,S] on the same device.
    The GPU body of dynamic_preprocess + build_transform; returns (tiles, tiles_per_frame)."""
    lib = _lib.load()
    if not frames_u8.is_cuda or frames_u8.dtype != torch.uint8 or frames_u8.dim() != 4 or frames_u8.shape[-1] != 3:
        raise TypeError("preprocess_frames expects a CUDA uint8 tensor [F,H,W,3]")
    frames_u8 = frames_u8.contiguous()
    F, H, W, _ = frames_u8.shape
    tpf = C.c_int32()
    out_dt = {torch.float32: _lib.MJV_DTYPE_F32, torch.bfloat16: _lib.MJV_DTYPE_BF16}[dtype]
    with torch.cuda.device(frames_u8.device):
        st = _lib.current_stream()
        _lib.check(lib.mjv_preprocess_frames(frames_u8.data_ptr(), F, H, W, input_size, max_num, int(use_thumbnail),
                                             None, out_dt, None, 0, C.byref(tpf), st))
        if out is None:
            out = torch.empty((F * tpf.value, 3, input_size, input_size), dtype=dtype, device=frames_u8.device)
        elif out.numel() != F * tpf.value * 3 * input_size * input_size or out.dtype != dtype or not out.is_contiguous():
            raise ValueError("preprocess_frames: `out` has the wrong size / dtype")
        # the fused resize kernel needs no scratch; only the two-pass fallback (a block that does not fit in shared memory,
        # or MJV_RESIZE_TWO_PASS=1) asks for the uint8 [F,H,OW,3] intermediate, by returning MJV_ERR_NOMEM
        rc = lib.mjv_preprocess_frames(frames_u8.data_ptr(), F, H, W, input_size, max_num, int(use_thumbnail),
                                       out.data_ptr(), out_dt, None, 0, C.byref(tpf), st)
        if rc == _lib.MJV_ERR_NOMEM:
            gi, gj = find_closest_aspect_ratio(W, H, 1, max_num, input_size)
            scratch = torch.empty(F * H * max(input_size * gi, input_size) * 3, dtype=torch.uint8, device=frames_u8.device)
            rc = lib.mjv_preprocess_frames(frames_u8.data_ptr(), F, H, W, input_size, max_num, int(use_thumbnail),
                                           out.data_ptr(), out_dt, scratch.data_ptr(), scratch.numel(), C.byref(tpf), st)
        _lib.check(rc)
    return out, tpf.value


def load_video(video_path, bound=None, input_size=448, max_num=1, num_segments=32, *, device=None, dtype=None):
    """Reference signature (data.py:158). Returns (pixel_values, num_patches_list).
    device=None -> CPU fp32 result like the reference; device='cuda' keeps the tiles on the GPU (optionally bf16)."""
    frames = sample_clip(video_path, bound, num_segments)
    frame_indices = range(frames.shape[0])
    if not torch.cuda.is_available():
        raise RuntimeError("load_video (mj_video_b200) preprocesses frames on the GPU; no CUDA device is available")
    gpu = torch.device(device if device is not None else "cuda")
    host = torch.from_numpy(frames)
    dev_frames = host.pin_memory().to(gpu, non_blocking=True) if host.numel() else host.to(gpu)
    tiles, tpf = preprocess_frames(dev_frames, input_size=input_size, max_num=max_num, use_thumbnail=True,
                                   dtype=dtype or torch.float32)
    num_patches_list = [tpf] * len(frame_indices)
    if device is None:
        tiles = tiles.cpu()
    return tiles, num_patches_list
