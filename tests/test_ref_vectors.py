"""Reference-side vectors (tools/ref_vectors/README.md): outputs of the real image 0.25.2 / kiddo 4.2.0 crates and of the
reference binary itself, consumed IF PRESENT under tests/golden/ref/.  They cannot be produced in the GPU image (no cargo,
no network), so today every case here is skipped with that reason; the day someone runs
`tools/ref_vectors/run_reference.sh <reference checkout>` these tests turn DESIGN.md §2's "parity unpinned" items into
checked ones — the CPU oracle under `-m "not gpu"`, the CUDA path under `-m gpu`."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("EMOSAIC_REF_VECTORS", os.path.join(ROOT, "tests", "golden", "ref"))
sys.path.insert(0, os.path.join(ROOT, "tools", "ref_vectors"))
import inputs as ref_inputs      # noqa: E402  tools/ref_vectors/inputs.py
import e2e_inputs                # noqa: E402  tools/ref_vectors/e2e_inputs.py

import oracle                    # noqa: E402


def ref(name):
    p = os.path.join(REF, name)
    if not os.path.exists(p):
        pytest.skip(f"{os.path.relpath(p, ROOT)} absent: reference-side vectors need cargo (tools/ref_vectors/README.md)")
    return np.load(p) if name.endswith(".npy") else p


def queries_as_source(q: np.ndarray) -> np.ndarray:
    """[Q, 3N] query vectors (cells row-major, analysis.rs:23-36) -> a [dim, Q*dim, 3] source image with one block per query."""
    Q, D = q.shape
    dim = int(round((D // 3) ** 0.5))
    return np.ascontiguousarray(q.reshape(Q, dim, dim, 3).transpose(1, 0, 2, 3).reshape(dim, Q * dim, 3))


def check_match(name, item, dist, colors):
    """item / dist of the path under test against kiddo's.  Distances must be identical; where the chosen tile differs the
    failure message says whether it is a tie-break difference (same distance, both in the argmin set) or a wrong answer."""
    ri, rd = ref(f"ref_kiddo_{name}_item.npy"), ref(f"ref_kiddo_{name}_dist.npy")
    item, dist = item.reshape(-1), dist.reshape(-1)
    assert (dist == rd).all(), f"{name}: {(dist != rd).sum()} distances differ from kiddo"
    if not (item == ri).all():
        bad = np.nonzero(item != ri)[0]
        raise AssertionError(f"{name}: {bad.size} of {item.size} queries pick another tile than kiddo at equal distance — the "
                             f"canonical tie-break (smallest index, unflipped first) is not kiddo's here; first: query {bad[0]} "
                             f"ours {item[bad[0]]} kiddo {ri[bad[0]]}")


# ---- CPU: the oracle against the crates -------------------------------------------------------------------------------
def test_oracle_blend_vs_image_crate():
    alphas, table = ref("ref_blend_alphas.npy"), ref("ref_blend.npy")
    for k, A in enumerate(alphas):
        lut, alpha_byte = oracle.blend_lut(int(A))
        assert (lut == table[k, :, :, 0]).all(), f"A={A}: {(lut != table[k, :, :, 0]).sum()} (bg, fg) pairs differ"
        assert (table[k, :, :, 1] == alpha_byte).all(), f"A={A}: output alpha byte"


def test_oracle_resize_nearest_vs_image_crate():
    cases, flat = ref("ref_resize_nearest_cases.npy"), ref("ref_resize_nearest.npy")
    for n_in, n_out, off in cases:
        ramp = np.zeros((1, n_in, 3), np.uint8)
        ramp[0, :, 0], ramp[0, :, 1] = np.arange(n_in) & 255, np.arange(n_in) >> 8
        out = oracle.tint(np.zeros((1, n_out, 3), np.uint8), ramp, 255)      # A = 255: the overlay pixel itself
        got = out[0, :, 0].astype(np.uint32) | out[0, :, 1].astype(np.uint32) << 8
        assert (got == flat[off:off + n_out]).all(), f"{n_in} -> {n_out}"


@pytest.mark.parametrize("name", sorted(ref_inputs.LANCZOS))
def test_oracle_lanczos3_vs_image_crate(name):
    want = ref(f"ref_lanczos3_{name}.npy")
    w, h, x0, y0, cw, ch, nw, nh = ref_inputs.LANCZOS[name]
    got = oracle.resize_lanczos3(ref_inputs.lanczos_input(name), nw, nh, (x0, y0, cw, ch))
    assert (got == want).all(), f"{name}: {(got != want).sum()} bytes differ, max |d| = {np.abs(got.astype(int) - want).max()}"


@pytest.mark.parametrize("name", sorted(ref_inputs.KIDDO))
def test_oracle_match_vs_kiddo(name):
    ref(f"ref_kiddo_{name}_item.npy")
    colors, q = ref_inputs.kiddo_input(name)
    T, Q, D = ref_inputs.KIDDO[name]
    item, dist = oracle.match(colors.reshape(T, D // 3, 3), queries_as_source(q))
    check_match(name, item, dist, colors)


@pytest.mark.parametrize("name", sorted(ref_inputs.KIDDO))
def test_oracle_candidate_order_vs_kiddo_nearest_n(name):
    """rendering.rs:307-321: the order of nearest_n among equal distances (the no-repeat lists)."""
    from oracle import oracle_np as onp
    ni, nd = ref(f"ref_kiddo_{name}_nearest_n_item.npy"), ref(f"ref_kiddo_{name}_nearest_n_dist.npy")
    colors, q = ref_inputs.kiddo_input(name)
    T, Q, D = ref_inputs.KIDDO[name]
    items, dists = onp.sorted_candidates(colors.reshape(T, D // 3, 3), queries_as_source(q[:ni.shape[0]]))
    k = ni.shape[1]
    assert (np.asarray(dists)[:, :k] == nd).all(), "distances of the candidate lists"
    assert (np.asarray(items)[:, :k] == ni).all(), "order among equal distances differs from kiddo's nearest_n"


def test_oracle_cache_bytes_vs_reference_binary():
    blob = open(ref("e2e_cache_1to1.bin"), "rb").read()
    from emosaic_b200 import cache
    colors, paths, dates = cache.deserialize_tile_set(blob, 1)
    assert len(paths) == e2e_inputs.T and all(p.startswith(e2e_inputs.WORK) for p in paths)
    assert cache.serialize_tile_set(colors, paths, dates) == blob               # same bytes back
    assert oracle.cache_serialize(colors, np.arange(1, len(paths) + 1), dates, paths) == blob


# ---- GPU: the CUDA path against the crates and the reference binary ------------------------------------------------------
@pytest.mark.gpu
def test_gpu_blend_and_nearest_vs_image_crate(ctx):
    alphas, table = ref("ref_blend_alphas.npy"), ref("ref_blend.npy")
    tiles = np.arange(256, dtype=np.uint8).reshape(256, 1, 1, 1).repeat(3, 3)    # tile k = one grey pixel of value k
    ctx.set_library(tiles.reshape(256, 1, 3), tiles)
    item = (np.arange(256, dtype=np.int32) + 1).reshape(256, 1).repeat(256, 1)   # row bg: tile of grey bg
    src = np.arange(256, dtype=np.uint8).reshape(1, 256, 1).repeat(256, 0).repeat(3, 2)  # column fg
    for k, A in enumerate(alphas):
        if A == 0:
            continue                                                             # tint 0 takes the RGB path (main.rs:447)
        out = ctx.compose(item, src, 4, int(A))
        assert (out[..., 0] == table[k, :, :, 0]).all() and (out[..., 3] == table[k, :, :, 1]).all(), f"A={A}"
    cases, flat = ref("ref_resize_nearest_cases.npy"), ref("ref_resize_nearest.npy")
    one = np.zeros((1, 1, 1, 3), np.uint8)
    ctx.set_library(one.reshape(1, 1, 3), one)
    for n_in, n_out, off in cases:
        ramp = np.zeros((1, n_in, 3), np.uint8)
        ramp[0, :, 0], ramp[0, :, 1] = np.arange(n_in) & 255, np.arange(n_in) >> 8
        out = ctx.compose_overlay(np.ones((1, n_out), np.int32), ramp, 255)
        got = out[0, :, 0].astype(np.uint32) | out[0, :, 1].astype(np.uint32) << 8
        assert (got == flat[off:off + n_out]).all(), f"{n_in} -> {n_out}"


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(ref_inputs.LANCZOS))
def test_gpu_lanczos3_vs_image_crate(ctx, name):
    want = ref(f"ref_lanczos3_{name}.npy")
    w, h, x0, y0, cw, ch, nw, nh = ref_inputs.LANCZOS[name]
    assert (ctx.resize(ref_inputs.lanczos_input(name), nw, nh, (x0, y0, cw, ch)) == want).all()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(ref_inputs.KIDDO))
def test_gpu_match_vs_kiddo(ctx, name):
    ref(f"ref_kiddo_{name}_item.npy")
    colors, q = ref_inputs.kiddo_input(name)
    T, Q, D = ref_inputs.KIDDO[name]
    ctx.set_library(colors.reshape(T, D // 3, 3))
    item, dist = ctx.match(queries_as_source(q))
    check_match(name, item, dist, colors)


@pytest.mark.gpu
@pytest.mark.parametrize("out,argv", [
    ("e2e_1to1.png", ["-s", "16", "--mode", "1"]), ("e2e_4to1.png", ["-s", "16", "--mode", "2"]),
    ("e2e_9to1.png", ["-s", "12", "--mode", "3"]), ("e2e_tint.png", ["-s", "16", "--mode", "1", "-t", "0.5"]),
    ("e2e_tint_a200.png", ["-s", "16", "--mode", "2", "-t", "0.7843137254901961"]),
    ("e2e_downsample.png", ["-s", "16", "--mode", "2", "--downsample", "3"]),
    ("e2e_no_repeat.png", ["-s", "8", "--mode", "1", "--no-repeat", "--downsample", "4"]),
])
def test_gpu_cli_vs_reference_binary(out, argv, tmp_path):
    """The whole command line against the image the reference binary wrote for the same tiles directory and source."""
    from PIL import Image
    from emosaic_b200 import cli
    want = np.asarray(Image.open(ref(out)))
    e2e_inputs.main()
    mine = tmp_path / out
    pre, post = argv[:2], argv[2:]
    assert cli.main(pre + ["-o", str(mine), os.path.join(e2e_inputs.WORK, "src.png"), "mosaic",
                           os.path.join(e2e_inputs.WORK, "tiles"), "--extensions", "png", "-f"] + post) == 0
    got = np.asarray(Image.open(mine))
    assert got.shape == want.shape
    assert (got == want).all(), f"{out}: {(got != want).any(-1).sum()} pixels differ from the reference binary's output"
    if "--mode" in argv and argv[argv.index("--mode") + 1] in ("1", "2") and "-t" not in argv and "--downsample" not in argv:
        n = {"1": 1, "2": 4}[argv[argv.index("--mode") + 1]]
        cache_ref = ref(f"e2e_cache_{n}to1.bin")
        assert open(os.path.join(e2e_inputs.WORK, "tiles", f".emosaic_{n}to1"), "rb").read() == open(cache_ref, "rb").read()
