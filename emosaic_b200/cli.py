"""Command line with the reference's surface for the accelerated path.

    python -m emosaic_b200 [-s N] [-o PATH] [--crop] IMG mosaic TILES_DIR [--mode 1|2|...|128|random] [-f] [-t X]
                           [--downsample K] [--extensions jpg jpeg]
    python -m emosaic_b200 ... IMG mosaic TILES_DIR -m 1to1|4to1|random        (README spelling)

Mirrors src/main.rs:28-138 (flags), :542-667 (n_to_1: dimension rule, cache, render) and :447-478 (tint).
Decoding (PIL) and directory walking stay on the host and outside the accelerated path.  Both Lanczos3 resizes of the
reference run on the GPU (emo_resize, bit-exact with image 0.25.2): the source image (main.rs:595) and tile preparation
(tiles/utils.rs:63-196: white-border trim view, optional centre-square crop, resize, EXIF rotation); the reference's
~/.cache/mosaic JPEG cache of prepared tiles is not reproduced (every run prepares from the original file).
--no-repeat runs the no-repeat renderer (rendering.rs:262-401); --randomize / --greedy / --html / --web are rejected:
outside the accelerated path.
"""
from __future__ import annotations

import argparse
import os
import sys
from typing import List, Optional

import numpy as np

from . import api, cache, stats

MODES = {"1": 1, "2": 2, "3": 3, "4": 4, "5": 5, "6": 6, "8": 8, "16": 16, "32": 32, "64": 64, "128": 128,
         "1to1": 1, "4to1": 2, "random": "random"}


def find_images(root: str, extensions) -> List[str]:
    """src/mosaic/image.rs:7-23: iterative DFS, extension predicate."""
    out, stack = [], [root]
    while stack:
        d = stack.pop()
        for e in sorted(os.scandir(d), key=lambda e: e.name):
            if e.is_dir(follow_symlinks=True):
                stack.append(e.path)
            elif os.path.splitext(e.name)[1][1:] in extensions:
                out.append(e.path)
    return out


def prepare_tiles(paths: List[str], tile_size: int, crop: bool, ctx, root: str = ""):
    """The tile loop of generate_tile_set (main.rs:757-806): a file that cannot be decoded or prepared (too small, no
    non-white interior, corrupt) is collected, listed and left out; the survivors keep their order, so idx numbering matches
    the reference.  Returns (pixels [T,ts,ts,3], surviving paths)."""
    px, ok, failed = [], [], []
    for p in paths:
        try:
            px.append(prepare_tile(p, tile_size, crop, ctx))
            ok.append(p)
        except (api.EmosaicError, OSError, ValueError, SyntaxError, EOFError) as e:   # PIL raises OSError / SyntaxError / ValueError
            failed.append((os.path.relpath(p, root) if root else p, e))
    if failed:
        print(f"Failed to read the following images({len(failed)}):", file=sys.stderr)
        for p, e in failed:
            print(f"- {p}: {e}", file=sys.stderr)
    return (np.stack(px) if px else np.zeros((0, tile_size, tile_size, 3), np.uint8)), ok


def prepare_tile(path: str, tile_size: int, crop: bool, ctx: Optional[api.Context] = None) -> np.ndarray:
    """tiles/utils.rs:63-196 without the JPEG cache: decode (host), then trim view / crop / Lanczos3 resize (GPU) / rotate."""
    from PIL import Image
    im = Image.open(path)
    try:
        orientation = int(im.getexif().get(274, 1))  # utils.rs:198-212 get_jpeg_orientation: 1..8, else 1
    except Exception:
        orientation = 1
    if not 1 <= orientation <= 8:
        orientation = 1
    return api.prepare_tile(np.asarray(im.convert("RGB"), dtype=np.uint8), tile_size, crop, orientation, ctx)


def exif_date(path: str) -> Optional[str]:
    try:
        from PIL import Image
        ex = Image.open(path).getexif()
        v = ex.get(36867) or ex.get_ifd(0x8769).get(36867) or ex.get(306)
        return str(v)[:10] if v else None
    except Exception:
        return None


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(prog="emosaic", description="Mosaic generator (B200 path)")
    p.add_argument("-s", "--tile-size", type=int, default=16)
    p.add_argument("-o", "--output-path", default="./output.jpg")
    p.add_argument("--crop", action="store_true")
    p.add_argument("--device", type=int, default=0)
    p.add_argument("--gpus", type=int, default=1, help="split the work over GPUs 0..N-1 of this box (row stripes of the source, "
                   "tile ranges of the analysis build; the library is replicated with one NCCL broadcast)")
    p.add_argument("img")
    sub = p.add_subparsers(dest="subcmd")
    sub.add_parser("prepare")
    m = sub.add_parser("mosaic")
    m.add_argument("tiles_dir")
    m.add_argument("-m", "--mode", default="1", choices=sorted(MODES))
    m.add_argument("-f", "--force", action="store_true")
    m.add_argument("-t", "--tint-opacity", type=float, default=0.0)
    m.add_argument("--no-repeat", action="store_true")
    m.add_argument("--downsample", type=int, default=1)
    m.add_argument("--randomize", type=float, default=None)
    m.add_argument("--extensions", nargs="+", default=["jpg", "jpeg"])
    m.add_argument("--greedy", action="store_true")
    m.add_argument("--html", action="store_true")
    m.add_argument("--web", action="store_true")
    m.add_argument("--title", default="Mosaic Widget")
    m.add_argument("--seed", type=int, default=None, help="random mode only: host RNG seed (the reference uses thread_rng)")
    return p


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    from PIL import Image
    ts = args.tile_size
    # main.rs:270-345 validate_tile_size / validate_input_image / validate_output_path
    if ts == 0:
        print("error: Tile size must be greater than 0", file=sys.stderr)
        return 1
    if ts > 1024:
        print("error: Tile size is too large (maximum: 1024)", file=sys.stderr)
        return 1
    if not os.path.isfile(args.img):
        print(f"error: Input image does not exist: {args.img}", file=sys.stderr)
        return 1
    if os.path.splitext(args.img)[1][1:].lower() not in ("jpg", "jpeg", "png", "bmp", "gif", "tiff", "webp"):
        print(f"error: Unsupported image format: {os.path.splitext(args.img)[1][1:]}", file=sys.stderr)
        return 1
    out_parent = os.path.dirname(args.output_path)
    if out_parent and not os.path.isdir(out_parent):
        print(f"error: Output directory does not exist: {out_parent}", file=sys.stderr)
        return 1
    if args.subcmd == "prepare":
        try:
            Image.fromarray(prepare_tile(args.img, ts, args.crop, api.Context(args.device))).save(args.output_path)
        except api.EmosaicError as e:
            print(f"error: {e}", file=sys.stderr)
            return 1
        return 0
    if args.subcmd != "mosaic":
        return 0  # main.rs:378-379 `None => ()`: without a subcommand the reference validates its arguments and does nothing
    if not (0.0 <= args.tint_opacity <= 1.0):
        print("error: Value must be between 0 and 1", file=sys.stderr)
        return 2
    if args.randomize is not None or args.greedy or args.html or args.web:
        # main.rs:663-667: --no-repeat alone is render_nto1_no_repeat (supported); with --greedy, or with --randomize,
        # it is render_nto1's rayon-order / thread_rng dependent branch
        print("error: --randomize/--greedy/--html/--web are outside the accelerated path", file=sys.stderr)
        return 2
    if not os.path.isdir(args.tiles_dir):
        print(f"error: tiles directory {args.tiles_dir} does not exist", file=sys.stderr)
        return 1
    mode = MODES[args.mode]
    print(f"Opening source image: {args.img}", file=sys.stderr)
    original = np.asarray(Image.open(args.img).convert("RGB"), dtype=np.uint8)
    if args.gpus < 1:
        print("error: --gpus must be at least 1", file=sys.stderr)
        return 2
    # one GPU: a Context; several: a Group (same interface for analyse_tiles / set_library / mosaic, sharded inside)
    ctx = api.Context(args.device) if args.gpus == 1 else api.Group(args.gpus)
    exts = set(args.extensions)

    if mode == "random":  # main.rs:414-442 + rendering.rs:418-440: uniform random tile per source pixel
        paths = [p for p in find_images(args.tiles_dir, exts) if os.path.exists(p)]
        print(f"Tile set with {len(paths)} tiles", file=sys.stderr)
        px, paths = prepare_tiles(paths, ts, True, ctx, args.tiles_dir)
        if len(paths) == 0:
            print("error: no tiles", file=sys.stderr)
            return 1
        ctx.set_library(np.zeros((len(paths), 1, 3), np.uint8), px)
        item = np.random.default_rng(args.seed).integers(1, len(paths) + 1, original.shape[:2]).astype(np.int32)
        out = ctx.compose(item)
        src_for_tint, dist, dim = original, None, 1
    else:
        dim = int(mode)
        N = dim * dim
        nw, nh = api.adjust_source_dims(original.shape[1], original.shape[0], args.downsample, dim)  # main.rs:567-587
        print(f"Resizing source image from {original.shape[1]}x{original.shape[0]} to {nw}x{nh}", file=sys.stderr)
        try:
            img = api.resize_source(original, args.downsample, dim, ctx)  # main.rs:595 imageops::resize(Lanczos3), on the GPU
        except api.EmosaicError as e:
            print(f"error: {e}", file=sys.stderr)
            return 1
        if img.shape[1] % dim or img.shape[0] % dim:
            print(f"Invalid source dimensions ({img.shape[1]}x{img.shape[0]}): Dimensions must be divisible by {dim}", file=sys.stderr)
            return 1
        if ts % dim:
            print(f"Invalid tile size: Tile size must be divisible by {dim}", file=sys.stderr)
            return 1
        cache_path = os.path.join(args.tiles_dir, cache.cache_file_name(N, args.crop))
        colors = paths = dates = px_render = None
        if not args.force and os.path.exists(cache_path):  # main.rs:617-654
            try:
                colors, paths, dates = cache.deserialize_tile_set(open(cache_path, "rb").read(), N, exts, check_exists=True)
                print("Reusing analysis cache", file=sys.stderr)
            except (ValueError, MemoryError, OverflowError):  # corrupt cache: re-analyse, like the reference's bincode Err
                colors = None
        if colors is None:  # generate_tile_set, main.rs:740-813, analysis on the GPU in one batch
            px, paths = prepare_tiles(find_images(args.tiles_dir, exts), ts, args.crop, ctx, args.tiles_dir)
            dates = [exif_date(p) for p in paths]
            colors = ctx.analyse_tiles(px, dim)
            with open(cache_path, "wb") as f:
                f.write(cache.serialize_tile_set(colors, paths, dates))
            if args.crop:
                px_render = px
        # tileset.rs:152-155: a TileSet built by from_tiles holds no images, so rendering always (re)prepares
        # the tiles with crop = true, whatever --crop was used for the analysis
        if px_render is None:
            px_render, kept = prepare_tiles(paths, ts, True, ctx, args.tiles_dir)
            if len(kept) != len(paths):  # tileset.rs:152-155 would fail at the first placement of such a tile (get_image -> Err)
                print("error: a tile of the analysed set could not be prepared for rendering", file=sys.stderr)
                return 1
        px = px_render
        print(f"Tile set with {len(paths)} tiles", file=sys.stderr)
        if len(paths) == 0:
            print("error: no tiles", file=sys.stderr)
            return 1
        if args.no_repeat:  # main.rs:663-664
            try:
                res = api.render_nto1_no_repeat(img, api.TileSet.from_arrays(colors, px, paths, dates), ts, ctx)
            except api.EmosaicError as e:
                print(f"error: {e}", file=sys.stderr)
                return 1
            out, item, dist = res.image, res.item, res.dist
            if args.tint_opacity > 0.0 and (item == 0).any():
                print("error: --tint-opacity with unplaced blocks (fewer tiles than blocks) is not supported", file=sys.stderr)
                return 1
        else:
            ctx.set_library(colors, px)
            out, item, dist = ctx.mosaic(img, 3, 0)
        placed = item != 0  # the reference records statistics for placed tiles only (rendering.rs:362-365)
        stats.summarise(item, dist, paths, ctx=ctx)   # counts and sums on the GPU (emo_stats), item 0 = unplaced: no entry
        src_for_tint = original  # main.rs:447-466 overlays the image as opened, not the copy resized for matching

    if args.tint_opacity > 0.0:  # main.rs:447-478: RGBA PNG, early return (no stats image)
        if src_for_tint.shape[0] == item.shape[0] * dim and src_for_tint.shape[1] == item.shape[1] * dim:
            rgba = ctx.compose(item, src_for_tint, 4, api.tint_alpha(args.tint_opacity))
        else:
            rgba = ctx.compose_overlay(item, src_for_tint, api.tint_alpha(args.tint_opacity))
        Image.fromarray(rgba, "RGBA").save(args.output_path, format="PNG")
        return 0
    print(f"Writing output file to {args.output_path}", file=sys.stderr)
    Image.fromarray(out).save(args.output_path, format="PNG")  # main.rs:483 save_with_format(Png)
    if dist is not None:
        sp = os.path.splitext(args.output_path)[0] + ".stats.png"
        # the no-repeat renderer keys its statistics by output coordinates (rendering.rs:352-365): one pixel per block
        Image.fromarray(stats.render(dist, ts if args.no_repeat else dim, ts, item != 0)).save(sp, format="PNG")
    return 0


if __name__ == "__main__":
    sys.exit(main())
