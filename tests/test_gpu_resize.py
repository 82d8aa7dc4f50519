"""emo_resize on the GPU against the CPU oracle: image 0.25.2 imageops::resize(view, nw, nh, Lanczos3), bit-exact
(main.rs:595 source resize; tiles/utils.rs:188-189 tile preparation).  Through the C ABI (ctypes)."""
import os

import numpy as np
import pytest

import oracle
from emosaic_b200 import api

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _photo(rng, h, w):
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([np.sin(xx / 11.0) * 110 + 128, np.cos(yy / 5.0) * 100 + 120, (xx * 3 + yy * 5) % 256], -1)
    return np.clip(base + rng.integers(-25, 26, base.shape), 0, 255).astype(np.uint8)


# (h, w, view or None, nh, nw): aligned rows (w % 4 == 0) and unaligned ones, views with odd offsets, up / down / near-identity,
# a single output pixel, one-pixel-wide and one-pixel-high inputs, a tail group narrower than 4 bytes
GEOMS = [
    (64, 64, None, 16, 16), (64, 64, None, 64, 64), (48, 52, None, 47, 51), (37, 41, None, 36, 40), (101, 103, None, 100, 102),
    (40, 52, (5, 3, 43, 33), 12, 12), (40, 52, (4, 0, 44, 40), 8, 8), (97, 131, (7, 5, 88, 88), 16, 16), (20, 24, None, 47, 61),
    (33, 29, (1, 2, 27, 30), 29, 27), (5, 5, None, 1, 1), (1, 1, None, 7, 3), (2, 300, None, 2, 11), (31, 2, None, 5, 2),
    (256, 256, None, 8, 8), (300, 200, None, 150, 100), (64, 67, None, 64, 66), (9, 7, None, 20, 31),
]


@pytest.mark.parametrize("h,w,view,nh,nw", GEOMS)
def test_resize_matches_oracle(ctx, h, w, view, nh, nw):
    rng = np.random.default_rng(h * 131 + w)
    for img in (rng.integers(0, 256, (h, w, 3), dtype=np.uint8), _photo(rng, h, w)):
        got = ctx.resize(img, nw, nh, view)
        assert got.shape == (nh, nw, 3)
        assert (got == oracle.resize_lanczos3(img, nw, nh, view)).all()


def test_resize_golden_vectors(ctx):
    g = np.load(os.path.join(GOLD, "resize_lanczos3.npz"))
    for k in range(int(g["cases"])):
        want = g[f"out{k}"]
        got = ctx.resize(g[f"img{k}"], want.shape[1], want.shape[0], tuple(int(v) for v in g[f"view{k}"]))
        assert (got == want).all()


def test_resize_batch_and_geometry_changes(ctx):
    """A batch shares one view; the cached tap tables must follow every change of geometry."""
    rng = np.random.default_rng(3)
    imgs = rng.integers(0, 256, (37, 50, 44, 3), dtype=np.uint8)
    for view, nw, nh in [(None, 8, 8), ((3, 1, 40, 47), 8, 8), ((3, 1, 40, 47), 9, 8), (None, 44, 50), (None, 8, 8)]:
        got = ctx.resize(imgs, nw, nh, view)
        want = np.stack([oracle.resize_lanczos3(im, nw, nh, view) for im in imgs])
        assert (got == want).all()


def test_resize_properties_large(ctx):
    """Full-size inputs: a flat 12 MP image stays flat (normalised weights), equal dimensions copy, a source-style 2050 -> 2048
    resize against the oracle, and a batch member does not depend on its neighbours or its position in the batch."""
    rng = np.random.default_rng(9)
    flat = np.empty((3001, 4003, 3), np.uint8)
    flat[:] = (13, 250, 128)
    out = ctx.resize(flat, 1000, 750)
    assert (out == (13, 250, 128)).all()
    big = rng.integers(0, 256, (2048, 2050, 3), dtype=np.uint8)
    assert (ctx.resize(big, 2050, 2048) == big).all()
    near = ctx.resize(big, 2048, 2048)                      # source-style: 2050 -> 2048 columns (main.rs:567-595)
    assert (near == oracle.resize_lanczos3(big, 2048, 2048)).all()   # the C oracle (OpenMP) still does this one in a second
    batch = rng.integers(0, 256, (9, 512, 512, 3), dtype=np.uint8)
    tiles = ctx.resize(batch, 64, 64)
    assert (tiles[4] == ctx.resize(batch[4], 64, 64)).all() and (tiles[8] == oracle.resize_lanczos3(batch[8], 64, 64)).all()


@pytest.mark.parametrize("n,h,w,view,ts", [(17, 512, 512, None, 64), (16, 515, 521, (3, 1, 512, 513), 64), (70, 300, 280, (0, 0, 280, 299), 32),
                                           (33, 512, 512, (1, 0, 511, 512), 48)])
def test_resize_tile_batches_transposed_path(ctx, n, h, w, view, ts):
    """Batches of photo -> tile reductions (>= 8x on both axes, >= 65536 outputs) take the transposed intermediate
    (resize.cu); aligned and unaligned rows, views, a tile size that is not a multiple of 32."""
    rng = np.random.default_rng(n * 7 + ts)
    imgs = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
    imgs[n // 2] = _photo(rng, h, w)
    got = ctx.resize(imgs, ts, ts, view)
    want = np.stack([oracle.resize_lanczos3(im, ts, ts, view) for im in imgs])
    assert (got == want).all()


def test_resize_full_grids_paired_rows_kernel(ctx):
    """Aligned geometries whose grids exceed one wave take the paired-rows vertical kernel in its 8-bytes-per-thread, one-group-
    ahead shape (resize.cu; smaller grids take 4 bytes and three groups): a batch through the transposed intermediate, a row-major
    down-scale with an odd number of output rows (the last pair has one row), an up-scale (windows of neighbouring rows coincide
    or leave gaps) and a view at an 8-byte aligned offset."""
    rng = np.random.default_rng(77)
    batch = rng.integers(0, 256, (80, 512, 512, 3), dtype=np.uint8)
    got = ctx.resize(batch, 64, 64)
    for k in (0, 41, 79):
        assert (got[k] == oracle.resize_lanczos3(batch[k], 64, 64)).all()
    big = rng.integers(0, 256, (2048, 2048, 3), dtype=np.uint8)
    big[300:900, 200:1500] = _photo(rng, 600, 1300)
    assert (ctx.resize(big, 1024, 1023) == oracle.resize_lanczos3(big, 1024, 1023)).all()
    view = (8, 3, 2024, 2040)
    assert (ctx.resize(big, 700, 701, view) == oracle.resize_lanczos3(big, 700, 701, view)).all()
    small = rng.integers(0, 256, (640, 640, 3), dtype=np.uint8)
    assert (ctx.resize(small, 640, 2049) == oracle.resize_lanczos3(small, 640, 2049)).all()


def test_resize_device_pointers_unaligned_and_guarded(ctx):
    """Device-pointer entry: odd base addresses (byte path), canaries around the output."""
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(21)
    for (n, h, w, view, nh, nw, shift) in [(3, 40, 52, None, 12, 12, 0), (3, 40, 52, None, 12, 12, 1), (2, 33, 29, (1, 2, 27, 30), 7, 9, 3),
                                           (1, 64, 64, None, 64, 64, 2), (5, 16, 16, None, 33, 31, 0)]:
        imgs = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
        src = torch.zeros(imgs.size + 16, dtype=torch.uint8, device=dev)
        src[shift:shift + imgs.size] = torch.from_numpy(imgs.reshape(-1)).to(dev)
        nbytes, pad = n * nh * nw * 3, 4096
        buf = torch.full((pad + nbytes + pad,), 0xA5, dtype=torch.uint8, device=dev)
        s = torch.cuda.current_stream().cuda_stream
        ctx.set_stream(s)
        try:
            ctx.resize_dev(src.data_ptr() + shift, n, w, h, view, nw, nh, buf.data_ptr() + pad)
            ctx.sync()
        finally:
            ctx.set_stream(None)
        got = buf[pad:pad + nbytes].cpu().numpy().reshape(n, nh, nw, 3)
        assert (got == np.stack([oracle.resize_lanczos3(im, nw, nh, view) for im in imgs])).all()
        assert bool((buf[:pad] == 0xA5).all()) and bool((buf[pad + nbytes:] == 0xA5).all())


def test_resize_rejects(ctx):
    img = np.zeros((10, 12, 3), np.uint8)
    with pytest.raises(api.EmosaicError, match="extends beyond"):
        ctx.resize(img, 4, 4, (5, 0, 8, 10))
    with pytest.raises(api.EmosaicError, match="empty"):
        ctx.resize(img, 0, 4)
    with pytest.raises(api.EmosaicError, match="empty"):
        ctx.resize(img, 4, 4, (0, 0, 0, 10))
    with pytest.raises(api.EmosaicError):
        ctx.resize(np.zeros((10, 12, 4), np.uint8), 4, 4)


def test_prepare_tile_and_source_resize(ctx):
    """prepare_tile (tiles/utils.rs:63-196) from the decoded photo on, and n_to_1's source resize (main.rs:567-595)."""
    rng = np.random.default_rng(17)
    photo = np.full((70, 90, 3), 255, np.uint8)
    photo[6:61, 9:84] = _photo(rng, 55, 75) // 2 + 20
    for crop in (False, True):
        for orientation in (1, 3, 6, 8, 5):
            t = api.prepare_tile(photo, 32, crop, orientation, ctx)
            assert t.shape == (32, 32, 3)                                      # utils.rs:291-299
            want = api.rotate(oracle.resize_lanczos3(photo, 32, 32, oracle.prepare_view(photo, 32, crop)), orientation)
            assert (t == want).all()
    src = _photo(rng, 101, 103)
    for downsample, dim in [(1, 2), (2, 3), (1, 1), (3, 4)]:
        got = api.resize_source(src, downsample, dim, ctx)
        nw, nh = oracle.adjust_dims(103, 101, downsample, dim)
        assert got.shape == (nh, nw, 3) and nw % dim == 0 and nh % dim == 0
        assert (got == oracle.resize_lanczos3(src, nw, nh)).all()


def test_cli_non_divisible_source_and_downsample(tmp_path, ctx):
    """The command line on a 37 x 45 source in 4to1 mode: the matched copy is the Lanczos3 resize to 36 x 44 (main.rs:567-595)."""
    PIL = pytest.importorskip("PIL.Image")
    from emosaic_b200 import cli
    rng = np.random.default_rng(23)
    tiles_dir = tmp_path / "tiles"
    tiles_dir.mkdir()
    for i in range(40):
        PIL.fromarray(np.minimum(_photo(rng, 30 + i % 7, 36 + i % 5), 230)).save(tiles_dir / f"t{i:02d}.png")
    src = _photo(rng, 37, 45)
    PIL.fromarray(src).save(tmp_path / "src.png")
    ts = 8
    paths = cli.find_images(str(tiles_dir), {"png"})

    def tile(p, crop):
        im = np.asarray(PIL.open(p).convert("RGB"), dtype=np.uint8)
        return oracle.resize_lanczos3(im, ts, ts, oracle.prepare_view(im, ts, crop))

    px_an, px_rd = np.stack([tile(p, False) for p in paths]), np.stack([tile(p, True) for p in paths])
    colors = oracle.analyse_tiles(px_an, 4)
    for downsample in (1, 2):
        out = tmp_path / f"o{downsample}.png"
        assert cli.main(["-s", str(ts), "-o", str(out), str(tmp_path / "src.png"), "mosaic", str(tiles_dir), "-m", "2", "-f",
                         "--extensions", "png", "--downsample", str(downsample)]) == 0
        nw, nh = oracle.adjust_dims(45, 37, downsample, 2)
        item, _ = oracle.match(colors, oracle.resize_lanczos3(src, nw, nh))
        assert (np.asarray(PIL.open(out)) == oracle.render(px_rd, item)).all()
