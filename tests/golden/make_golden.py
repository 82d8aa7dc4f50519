"""Generates tests/golden/*.npz.

The reference is a Rust crate that cannot be built or imported in this image (no cargo/rustc,
nightly-only, 214 un-vendored crates), so these vectors are NOT outputs of the reference binary.
They are: (a) the known-answer vectors copied from the reference's own unit tests (file:line in
`kat` below) and (b) seeded inputs with the outputs of the CPU oracle (oracle/emosaic_oracle.c),
which itself is pinned by (a).  Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def case(name, N, T, ts, H, W, seed, A=None, smooth=False):
    rng = np.random.default_rng(seed)
    tiles = rng.integers(0, 256, (T, ts, ts, 3), dtype=np.uint8)
    if smooth:
        yy, xx = np.mgrid[0:H, 0:W]
        src = np.stack([(xx * 255 // max(W - 1, 1)), (yy * 255 // max(H - 1, 1)), ((xx + yy) * 255 // max(W + H - 2, 1))], -1)
        src = np.clip(src + rng.integers(-6, 7, src.shape), 0, 255).astype(np.uint8)
    else:
        src = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    colors = oracle.analyse_tiles(tiles, N)
    item, dist = oracle.match(colors, src)
    out = oracle.render(tiles, item)
    # composited images are stored as SHA-256 digests (random pixels do not compress)
    d = dict(N=N, tiles=tiles, src=src, colors=colors, item=item, dist=dist, out_shape=np.array(out.shape),
             out_sha256=sha(out))
    if A is not None:
        t = oracle.tint(out, src, A)
        d["A"] = A
        d["tint_shape"] = np.array(t.shape)
        d["tint_sha256"] = sha(t)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    print(name, "T", T, "ts", ts, "src", src.shape, "flipped", int((item < 0).sum()), "maxdist", int(dist.max()))


if __name__ == "__main__":
    # C1-shaped: 1to1, ts 16, T=300 (single kiddo leaf: tie-break provable), 100x100 source (smaller ts to keep the file small)
    case("c1_1to1_t300", 1, 300, 16, 100, 100, 1234)
    case("c1_1to1_t300_smooth", 1, 300, 16, 60, 60, 1235, smooth=True)
    # 4to1 with mirrored winners, ts 16
    case("c2_4to1_small", 4, 500, 16, 64, 96, 1236)
    # 9to1 (mode 3), ts 12
    case("m3_9to1_small", 9, 200, 12, 30, 42, 1237)
    # tint (A=127 = opacity 0.5) 1to1 ts 8 and 4to1 ts 16
    case("c5_tint_1to1", 1, 128, 8, 40, 56, 1238, A=127)
    case("tint_4to1_a200", 4, 64, 16, 24, 32, 1239, A=200)
    # few-colour library: lots of exact ties (duplicate tiles)
    rng = np.random.default_rng(99)
    T, ts = 64, 8
    pal = rng.integers(0, 256, (8, 3), dtype=np.uint8)
    tiles = np.repeat(pal[rng.integers(0, 8, T)][:, None, None, :], ts, 1).repeat(ts, 2).astype(np.uint8)
    src = pal[rng.integers(0, 8, (32, 32))].astype(np.uint8)
    colors = oracle.analyse_tiles(tiles, 1)
    item, dist = oracle.match(colors, src)
    np.savez_compressed(os.path.join(HERE, "ties_palette.npz"), N=1, tiles=tiles, src=src, colors=colors, item=item,
                        dist=dist, out_sha256=sha(oracle.render(tiles, item)))
    print("ties_palette", int(dist.max()), np.unique(item).size)
    # .emosaic_4to1 bytes (bincode 1.3.3 defaults) for 5 tiles: dates on tiles 1 and 3, one non-ASCII path
    rng = np.random.default_rng(7)
    ccol = rng.integers(0, 256, (5, 4, 3), dtype=np.uint8)
    cpaths = ["/photos/a.jpg", "/photos/sub dir/b.jpeg", "/photos/c.jpg", "/photos/déjà.jpg", "rel/e.jpg"]
    cdates = [None, "2021:06:01", None, "2019:12:31", None]
    blob = oracle.cache_serialize(ccol, np.arange(1, 6), cdates, cpaths)
    open(os.path.join(HERE, "cache_4to1.bin"), "wb").write(blob)
    np.savez_compressed(os.path.join(HERE, "cache_4to1_meta.npz"), colors=ccol, paths=np.array(cpaths), dates=np.array([d or "" for d in cdates]))
    print("cache_4to1.bin", len(blob), "bytes")
    # Lanczos3 resize (image 0.25.2 imageops::resize; main.rs:595, tiles/utils.rs:188-189): source-style near-identity
    # resizes, a downsample by 2, a photo -> tile reduction with a view, an upscale
    rng = np.random.default_rng(2024)
    yy, xx = np.mgrid[0:97, 0:131]
    photo = np.stack([(np.sin(xx / 9.0) * 100 + 128), (np.cos(yy / 7.0) * 90 + 120), ((xx * yy) % 256)], -1)
    photo = np.clip(photo + rng.integers(-20, 21, photo.shape), 0, 255).astype(np.uint8)
    rcases = [(photo, (0, 0, 131, 97), 130, 96), (photo, (0, 0, 131, 97), 65, 48), (photo, (7, 5, 88, 88), 16, 16),
              (photo[:20, :24], (0, 0, 24, 20), 61, 47), (rng.integers(0, 256, (33, 29, 3), dtype=np.uint8), (1, 2, 27, 30), 27, 29)]
    d = {"cases": len(rcases)}
    for k, (img, view, nw, nh) in enumerate(rcases):
        d[f"img{k}"], d[f"view{k}"], d[f"out{k}"] = img, np.array(view), oracle.resize_lanczos3(img, nw, nh, view)
    np.savez_compressed(os.path.join(HERE, "resize_lanczos3.npz"), **d)
    print("resize_lanczos3", len(rcases), "cases")
