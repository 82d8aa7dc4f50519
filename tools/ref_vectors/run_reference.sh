#!/bin/bash
# End-to-end vectors from the REAL reference binary (needs rustup with the nightly named in the reference's
# rust-toolchain.toml, and network access for its crates; neither exists in the GPU image).
#   usage: tools/ref_vectors/run_reference.sh <checkout of pepeiborra/emosaic> [<output dir, default tests/golden/ref>]
# What it pins that nothing else can: the whole path of `emosaic IMG mosaic TILES` — prepare_tile (trim, Lanczos3), the
# analysis cache bytes (bincode), kiddo's choice among equidistant tiles for 600 tiles (above its single-leaf regime),
# the mirrored matches of 4to1 and the tinted RGBA output — as PNGs and cache files that tests/test_ref_vectors.py replays
# through this repository's command line.
set -euo pipefail
REF=$(realpath "$1")
OUT=$(realpath "${2:-$(dirname "$0")/../../tests/golden/ref}")
HERE=$(dirname "$(realpath "$0")")
WORK=/tmp/emosaic_ref_e2e
mkdir -p "$OUT"
python "$HERE/e2e_inputs.py"
(cd "$REF" && cargo build --release)
BIN="$REF/target/release/emosaic"
# a fresh HOME: the reference caches prepared tiles as JPEG under ~/.cache/mosaic and would reuse the lossy copies
# (tiles/utils.rs:73-85); without --crop the render phase prepares every tile from its source file (tileset.rs:152-155)
run() { HOME=$(mktemp -d) "$BIN" "$@"; }
run -s 16 -o "$OUT/e2e_1to1.png"  "$WORK/src.png" mosaic "$WORK/tiles" --mode 1 --extensions png -f
cp "$WORK/tiles/.emosaic_1to1" "$OUT/e2e_cache_1to1.bin"
run -s 16 -o "$OUT/e2e_4to1.png"  "$WORK/src.png" mosaic "$WORK/tiles" --mode 2 --extensions png -f
cp "$WORK/tiles/.emosaic_4to1" "$OUT/e2e_cache_4to1.bin"
run -s 12 -o "$OUT/e2e_9to1.png"  "$WORK/src.png" mosaic "$WORK/tiles" --mode 3 --extensions png -f
run -s 16 -o "$OUT/e2e_tint.png"  "$WORK/src.png" mosaic "$WORK/tiles" --mode 1 --extensions png -f -t 0.5
run -s 16 -o "$OUT/e2e_tint_a200.png" "$WORK/src.png" mosaic "$WORK/tiles" --mode 2 --extensions png -f -t 0.7843137254901961
run -s 16 -o "$OUT/e2e_downsample.png" "$WORK/src.png" mosaic "$WORK/tiles" --mode 2 --extensions png -f --downsample 3
run -s 8  -o "$OUT/e2e_no_repeat.png" "$WORK/src.png" mosaic "$WORK/tiles" --mode 1 --extensions png -f --no-repeat --downsample 4
# crate-level vectors (stable Rust is enough)
python "$HERE/inputs.py" "$HERE/inputs"
cargo run --release --manifest-path "$HERE/Cargo.toml" -- "$HERE/inputs" "$OUT"
ls -la "$OUT"
