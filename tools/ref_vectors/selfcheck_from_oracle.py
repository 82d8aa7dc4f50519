"""Plumbing check for tests/test_ref_vectors.py: writes the crate-level vector files FROM THE CPU ORACLE into a scratch
directory, so the consuming tests can be exercised without cargo:

    python tools/ref_vectors/selfcheck_from_oracle.py /tmp/ref_standin
    EMOSAIC_REF_VECTORS=/tmp/ref_standin python -m pytest tests/test_ref_vectors.py -q -m "not gpu"

These are NOT reference vectors (they can only agree with the oracle) and must never be copied to tests/golden/ref/."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import inputs  # noqa: E402
import oracle  # noqa: E402
from oracle import oracle_np as onp  # noqa: E402

out = sys.argv[1]
assert "golden" not in os.path.abspath(out), "scratch directories only"
os.makedirs(out, exist_ok=True)
alphas = np.array([1, 64, 127, 128, 200, 254, 0, 255], np.uint8)
table = np.zeros((len(alphas), 256, 256, 2), np.uint8)
for k, A in enumerate(alphas):
    lut, a = oracle.blend_lut(int(A))
    table[k, :, :, 0], table[k, :, :, 1] = lut, a
np.save(os.path.join(out, "ref_blend_alphas.npy"), alphas)
np.save(os.path.join(out, "ref_blend.npy"), table)
cases, flat = [], []
for n_in, n_out in [(100, 1600), (100, 800), (1024, 32768), (37, 296), (4097, 32768), (50, 75), (640, 481), (7, 1000)]:
    X = np.arange(n_out, dtype=np.float32)
    idx = np.minimum(np.floor((X + np.float32(0.5)) * (np.float32(n_in) / np.float32(n_out))).astype(np.uint32), n_in - 1)
    cases.append((n_in, n_out, sum(len(f) for f in flat)))
    flat.append(idx)
np.save(os.path.join(out, "ref_resize_nearest_cases.npy"), np.array(cases, np.uint32))
np.save(os.path.join(out, "ref_resize_nearest.npy"), np.concatenate(flat).astype(np.uint32))
for name, (w, h, x0, y0, cw, ch, nw, nh) in inputs.LANCZOS.items():
    np.save(os.path.join(out, f"ref_lanczos3_{name}.npy"), oracle.resize_lanczos3(inputs.lanczos_input(name), nw, nh, (x0, y0, cw, ch)))
for name, (T, Q, D) in inputs.KIDDO.items():
    colors, q = inputs.kiddo_input(name)
    dim = int(round((D // 3) ** 0.5))
    src = np.ascontiguousarray(q.reshape(Q, dim, dim, 3).transpose(1, 0, 2, 3).reshape(dim, Q * dim, 3))
    item, dist = oracle.match(colors.reshape(T, D // 3, 3), src)
    np.save(os.path.join(out, f"ref_kiddo_{name}_item.npy"), item.reshape(-1).astype(np.int32))
    np.save(os.path.join(out, f"ref_kiddo_{name}_dist.npy"), dist.reshape(-1).astype(np.uint32))
    qn = min(Q, 256)
    items, dists = onp.sorted_candidates(colors.reshape(T, D // 3, 3), src[:, :qn * dim])
    np.save(os.path.join(out, f"ref_kiddo_{name}_nearest_n_item.npy"), np.asarray(items)[:, :64].astype(np.int32))
    np.save(os.path.join(out, f"ref_kiddo_{name}_nearest_n_dist.npy"), np.asarray(dists)[:, :64].astype(np.uint32))
print("stand-in vectors (oracle-derived, plumbing check only) in", out)
