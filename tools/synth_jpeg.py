"""Deterministic synthetic JPEGs for the tests and the benchmark.

Pixel recipe (seeded): 1/16-resolution uniform noise bilinearly upsampled + one random-frequency
sinusoid of amplitude 24 + Gaussian noise sigma 4, clipped to u8.  Encoder: Pillow/libjpeg-turbo,
quality 85; even seeds use the Annex-K Huffman tables, odd seeds `optimize=True` (custom tables).
Parity never depends on the encoder: the GPU path and the oracle decode the same bytes.
"""
from __future__ import annotations

import io
import os
from concurrent.futures import ProcessPoolExecutor

import numpy as np


def synth_pixels(seed: int, w: int, h: int, c: int) -> np.ndarray:
    import cv2

    rng = np.random.default_rng(seed)
    lo = rng.integers(0, 256, (h // 16 + 2, w // 16 + 2, c)).astype(np.float32)
    fx, fy = rng.uniform(0.01, 0.08), rng.uniform(0.01, 0.08)
    up = cv2.resize(lo, ((w // 16 + 2) * 16, (h // 16 + 2) * 16), interpolation=cv2.INTER_LINEAR)
    if up.ndim == 2:
        up = up[:, :, None]
    img = up[8:8 + h, 8:8 + w, :]
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = img + (24.0 * np.sin(2 * np.pi * (xx * fx + yy * fy)))[:, :, None]
    img = img + 4.0 * rng.standard_normal((h, w, c), dtype=np.float32)
    return np.clip(img, 0, 255).astype(np.uint8)


def encode(seed: int, w: int, h: int, mode: str = "YCbCr", subsampling: str = "4:2:0", restart_rows: int = 0,
           restart_blocks: int = 0, progressive: bool = False, quality: int = 85, ycck: bool = False) -> bytes:
    """mode: 'L' (gray), 'YCbCr' (from RGB pixels), 'CMYK'."""
    from PIL import Image

    c = {"L": 1, "YCbCr": 3, "CMYK": 4}[mode]
    px = synth_pixels(seed, w, h, c)
    if mode == "L":
        im = Image.fromarray(px[:, :, 0], "L")
    elif mode == "YCbCr":
        im = Image.fromarray(px, "RGB")
    else:
        im = Image.fromarray(px, "CMYK")
    kw = dict(quality=quality, optimize=bool(seed & 1), progressive=progressive)
    if mode == "YCbCr":
        kw["subsampling"] = subsampling
    if restart_rows:
        kw["restart_marker_rows"] = restart_rows
    if restart_blocks:
        kw["restart_marker_blocks"] = restart_blocks
    buf = io.BytesIO()
    im.save(buf, "JPEG", **kw)
    data = buf.getvalue()
    if ycck:
        # Pillow cannot emit YCbCrK: flip the Adobe APP14 transform byte 0 -> 2
        i = data.index(b"\xff\xee")
        assert data[i + 4:i + 9] == b"Adobe"
        data = data[:i + 15] + b"\x02" + data[i + 16:]
    return data


def _job(args):
    return encode(*args[0], **args[1])


def make_batch(cfg: int, count: int, w: int, h: int, workers: int | None = None, cache_dir: str | None = None,
               first: int = 0, **kw) -> list[bytes]:
    """`count` images with seeds cfg*10000 + first + i.  Cached on disk when cache_dir is given."""
    key = f"cfg{cfg}_{w}x{h}_" + "_".join(f"{k}-{v}" for k, v in sorted(kw.items()))
    out: list[bytes | None] = [None] * count
    todo = []
    for i in range(count):
        if cache_dir:
            p = os.path.join(cache_dir, f"{key}_{first + i}.jpg")
            if os.path.exists(p):
                with open(p, "rb") as f:
                    out[i] = f.read()
                continue
        todo.append(i)
    if todo:
        jobs = [((cfg * 10000 + first + i, w, h), kw) for i in todo]
        workers = workers or min(len(todo), os.cpu_count() or 1)
        if workers > 1 and len(todo) > 1:
            with ProcessPoolExecutor(max_workers=workers) as ex:
                res = list(ex.map(_job, jobs, chunksize=max(1, len(jobs) // (workers * 4))))
        else:
            res = [_job(j) for j in jobs]
        for i, d in zip(todo, res):
            out[i] = d
            if cache_dir:
                os.makedirs(cache_dir, exist_ok=True)
                with open(os.path.join(cache_dir, f"{key}_{first + i}.jpg"), "wb") as f:
                    f.write(d)
    return out  # type: ignore
