"""Seeded inputs of the reference-vector recipe (tools/ref_vectors/README.md).  The Rust program reads them as raw bytes; the
tests regenerate the same arrays in memory to compute what the oracle and the CUDA path say.

    python tools/ref_vectors/inputs.py <dir>      # writes <dir>/manifest.txt + *.bin
"""
from __future__ import annotations

import os
import sys

import numpy as np

# name -> (w, h, view x0, y0, cw, ch, nw, nh): source-style and photo -> tile resizes (main.rs:595, tiles/utils.rs:188-189)
LANCZOS = {
    "near_identity": (130, 97, 0, 0, 130, 97, 128, 96),
    "half": (256, 200, 0, 0, 256, 200, 128, 100),
    "photo_to_tile": (400, 300, 50, 0, 300, 300, 16, 16),
    "view_odd": (211, 157, 5, 3, 199, 151, 64, 48),
    "upscale": (40, 30, 0, 0, 40, 30, 100, 75),
    "smooth": (320, 240, 0, 0, 320, 240, 32, 24),
}
# name -> (T, Q, D): libraries below and above kiddo's 640-entry bucket (T <= 320: one leaf), 1to1 / 4to1 / 9to1 vectors,
# quantised colours (many exact distance ties) and duplicated tiles
KIDDO = {
    "n1_t300": (300, 4096, 3),
    "n1_t2000": (2000, 4096, 3),
    "n1_t20000_quant": (20000, 4096, 3),
    "n4_t300": (300, 2048, 12),
    "n4_t5000": (5000, 2048, 12),
    "n4_t10000_quant": (10000, 2048, 12),
    "n9_t3000": (3000, 1024, 27),
}


def lanczos_input(name: str) -> np.ndarray:
    w, h = LANCZOS[name][:2]
    rng = np.random.default_rng(sum(map(ord, name)))
    if name == "smooth":
        yy, xx = np.mgrid[0:h, 0:w]
        img = np.stack([xx * 255 // (w - 1), yy * 255 // (h - 1), (xx + yy) * 255 // (w + h - 2)], -1)
        return np.clip(img + rng.integers(-6, 7, img.shape), 0, 255).astype(np.uint8)
    return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)


def kiddo_input(name: str):
    T, Q, D = KIDDO[name]
    rng = np.random.default_rng(sum(map(ord, name)) + 7)
    colors = rng.integers(0, 256, (T, D), dtype=np.uint8)
    queries = rng.integers(0, 256, (Q, D), dtype=np.uint8)
    if "quant" in name:              # coarse colours: exact ties between different tiles on most queries
        colors = (colors // 32 * 32).astype(np.uint8)
        queries = (queries // 16 * 16).astype(np.uint8)
    colors[T // 2:T // 2 + 5] = colors[:5]      # duplicated tiles: the smaller index must win (insertion order)
    queries[:5] = colors[:5]                     # exact hits
    return colors, queries


def main(out: str):
    os.makedirs(out, exist_ok=True)
    lines = []
    for name, (w, h, x0, y0, cw, ch, nw, nh) in LANCZOS.items():
        lanczos_input(name).tofile(os.path.join(out, f"lanczos_{name}.bin"))
        lines.append(f"lanczos {name} {w} {h} {x0} {y0} {cw} {ch} {nw} {nh}")
    for name, (T, Q, D) in KIDDO.items():
        c, q = kiddo_input(name)
        c.tofile(os.path.join(out, f"kiddo_colors_{name}.bin"))
        q.tofile(os.path.join(out, f"kiddo_queries_{name}.bin"))
        lines.append(f"kiddo {name} {T} {Q} {D}")
    open(os.path.join(out, "manifest.txt"), "w").write("\n".join(lines) + "\n")
    print(f"wrote {len(lines)} cases to {out}")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "inputs"))
