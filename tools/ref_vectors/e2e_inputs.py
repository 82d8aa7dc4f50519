"""Inputs of the end-to-end half of the recipe: a tiles directory of PNG photos and a source PNG at a FIXED location
(the analysis cache stores the tile paths as found, so the recipe and the test must use the same directory).

    python tools/ref_vectors/e2e_inputs.py            # writes /tmp/emosaic_ref_e2e/{tiles/*.png,src.png}
"""
from __future__ import annotations

import os

import numpy as np

WORK = "/tmp/emosaic_ref_e2e"
T = 600            # above kiddo's single-leaf regime (2T = 1200 > 640): exercises the tie-break parity left unpinned
TILE_SHAPE = (40, 52)
SRC_SHAPE = (60, 84)


def tiles_and_source():
    rng = np.random.default_rng(2024)
    tiles = []
    for i in range(T):
        base = rng.integers(0, 256, 3) // 16 * 16          # quantised base colours: exact ties between tiles
        img = np.clip(base + rng.integers(-3, 4, TILE_SHAPE + (3,)), 0, 235).astype(np.uint8)
        img[:, :20] = np.clip(img[:, :20].astype(int) - 40, 0, 255)   # left/right asymmetry: mirrored matches in 4to1
        tiles.append(img)
    src = (rng.integers(0, 256, SRC_SHAPE + (3,)) // 8 * 8).astype(np.uint8)
    return tiles, src


def main():
    from PIL import Image
    tiles, src = tiles_and_source()
    os.makedirs(os.path.join(WORK, "tiles"), exist_ok=True)
    for i, t in enumerate(tiles):
        Image.fromarray(t).save(os.path.join(WORK, "tiles", f"t{i:04d}.png"))
    Image.fromarray(src).save(os.path.join(WORK, "src.png"))
    print(f"wrote {len(tiles)} tiles and src.png under {WORK}")


if __name__ == "__main__":
    main()
