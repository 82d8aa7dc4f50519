/*
 * emosaic_cuda.h — C ABI of libemosaic_cuda.so, the B200 (sm_100a) implementation of
 * emosaic's data-parallel core: tile analysis -> L1 nearest-colour match -> compose/tint.
 *
 * The reference (pepeiborra/emosaic, Rust) has no FFI seam; the three in-process calls this
 * library replaces are cited on each entry point (paths relative to the reference root).
 * Entry points are batch-granular: one call covers the whole tile library / source stripe,
 * because a per-block call (the reference's closure granularity) cannot feed a GPU.
 *
 * Conventions
 *  - plain C types only; every function returns 0 on success or a negative emo_status;
 *    the message of the last failure on the calling thread is emo_last_error().
 *    Nothing unwinds across the boundary (the reference's panics/exit(1) become codes).
 *  - one emo_ctx per GPU; a ctx is driven by one host thread at a time.  Several GPUs are used either as one
 *    process per GPU (emo_comm_init_rank joins the ctx to an NCCL communicator) or from one process through an
 *    emo_group (one ctx per device, one worker thread per device inside the group calls); see "multi-GPU".
 *  - `*_dev` variants take DEVICE pointers (same CUDA primary context, e.g. memory from
 *    emo_dev_alloc or any CUDA allocator) and are asynchronous on the ctx stream;
 *    the un-suffixed variants take HOST pointers, stage through the ctx's own device
 *    buffers and are synchronous on return.
 *  - images are row-major interleaved u8: RGB8 [H,W,3] or RGBA8 [H,W,4].
 *  - item maps are int32, signed and 1-based like the reference's i16 ids
 *    (tileset.rs:131-143): +k = tile k, -k = tile k mirrored horizontally, 0 never occurs.
 *  - there is NO CPU fallback: without a usable CUDA device emo_create fails.
 */
#ifndef EMOSAIC_CUDA_H
#define EMOSAIC_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EMO_ABI_VERSION 2

typedef struct emo_ctx emo_ctx;

typedef enum emo_status {
    EMO_OK = 0,
    EMO_ERR_ARG = -1,         /* bad argument (null pointer, size rule violated, item out of range) */
    EMO_ERR_CUDA = -2,        /* a CUDA runtime call failed */
    EMO_ERR_OOM = -3,         /* device or pinned-host allocation failed */
    EMO_ERR_STATE = -4,       /* call sequence error (e.g. match before set_library) */
    EMO_ERR_UNSUPPORTED = -5, /* valid in the reference but not implemented by this build */
    EMO_ERR_NO_DEVICE = -6,   /* no CUDA device / wrong architecture: there is no CPU path */
    EMO_ERR_NCCL = -7         /* NCCL could not be loaded or a collective failed */
} emo_status;

int emo_abi_version(void);
const char *emo_last_error(void);

/* ---- context -------------------------------------------------------------------------- */
/* Creates the per-GPU context (stream, scratch buffers). device = CUDA ordinal. */
int emo_create(int device, emo_ctx **out);
void emo_destroy(emo_ctx *ctx);
/* Use an externally owned cudaStream_t (e.g. torch's current stream) for all work; NULL
 * restores the ctx's own stream.  The ctx's scratch buffers are ordered on the stream in use:
 * call emo_sync before switching streams while *_dev work is still in flight. */
int emo_set_stream(emo_ctx *ctx, void *cuda_stream);
int emo_sync(emo_ctx *ctx);
/* Device name / SM count / compute capability of the ctx's GPU. */
int emo_device_info(emo_ctx *ctx, char *name, size_t name_cap, int *sm_count, int *cc_major, int *cc_minor);
/* Number of kernel launches issued by this ctx since creation (bench bookkeeping). */
uint64_t emo_launch_count(emo_ctx *ctx);

/* CUDA-event stopwatch on the ctx's current stream. */
int emo_timer_start(emo_ctx *ctx);
int emo_timer_stop(emo_ctx *ctx, float *elapsed_ms); /* synchronises the stop event */
/* Numbered event marks on the ctx's current stream (slot < 65536), for per-kernel timing inside a
 * longer timed region: emo_mark records, emo_mark_elapsed waits for mark b and returns b - a. */
int emo_mark(emo_ctx *ctx, uint32_t slot);
int emo_mark_elapsed(emo_ctx *ctx, uint32_t a, uint32_t b, float *elapsed_ms);

/* Raw memory helpers so a host without its own CUDA bindings can keep data resident. */
int emo_dev_alloc(emo_ctx *ctx, size_t bytes, void **out);
int emo_dev_free(emo_ctx *ctx, void *p);
int emo_host_alloc(emo_ctx *ctx, size_t bytes, void **out); /* pinned */
int emo_host_free(emo_ctx *ctx, void *p);
int emo_copy_h2d(emo_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes); /* async */
int emo_copy_d2h(emo_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes); /* async */

/* ---- (0) Lanczos3 resize --------------------------------------------------------------------
 * Replaces image 0.25.2 imageops::resize(view, nw, nh, FilterType::Lanczos3) at its two call sites
 * on the path: the source image before matching (src/main.rs:595, after the dimension rule
 * main.rs:567-587) and the photo -> tile resize of prepare_tile (src/mosaic/tiles/utils.rs:188-189,
 * on the view left by the white-border trim / centre-square crop, utils.rs:93-186).
 * images [n, img_h, img_w, 3]; every image of the batch is cropped to the view (x0, y0, cw, ch) and
 * resized to out [n, nh, nw, 3].  Bit-exact restatement of the crate's two-pass f32 algorithm
 * (vertical_sample into f32, horizontal_sample with clamp + round half away from zero, every tap one
 * f32 multiply and one f32 add in ascending order); equal dimensions copy, as resize() does.  The tap
 * weights (O(nw + nh) values, libm sinf like f32::sin) are prepared on the host side of the library,
 * all per-pixel arithmetic runs on the GPU.  Limits: rows <= 65535. */
int emo_resize(emo_ctx *ctx, const uint8_t *images, uint32_t n, uint32_t img_w, uint32_t img_h, uint32_t x0,
               uint32_t y0, uint32_t cw, uint32_t ch, uint32_t nw, uint32_t nh, uint8_t *out);
int emo_resize_dev(emo_ctx *ctx, const uint8_t *images_dev, uint32_t n, uint32_t img_w, uint32_t img_h, uint32_t x0,
                   uint32_t y0, uint32_t cw, uint32_t ch, uint32_t nw, uint32_t nh, uint8_t *out_dev);

/* The tap table of one axis exactly as emo_resize computes it: for every output index the first source index
 * (left[n_out]), the tap count (cnt[n_out]) and the normalised Lanczos3 weights (ws[n_out][pitch], zero beyond
 * cnt).  Host-only (no ctx, no GPU): for inspection and for checking the host side of the library against
 * another implementation of image 0.25.2's sampler set-up.  Any output pointer may be NULL; *max_taps receives
 * the largest tap count (the minimum pitch). */
int emo_resize_taps(uint32_t n_in, uint32_t n_out, uint32_t *left, uint32_t *cnt, float *ws, uint32_t pitch,
                    uint32_t *max_taps);

/* ---- (1) tile analysis ------------------------------------------------------------------
 * Replaces analyse::<N>() (src/mosaic/analysis.rs:5-20) + average_color()
 * (src/mosaic/color.rs:14-42) looped over the library (src/main.rs:786-794).
 * tiles [T,ts,ts,3] -> out [T,dim*dim,3]; cell = floor(ts/dim) square, cells row-major,
 * per channel floor(sum/count) in integer arithmetic (bit-exact). dim=1 is "1to1",
 * dim=2 is "4to1".  Requires 1 <= dim <= ts <= 4096. */
int emo_analyse(emo_ctx *ctx, const uint8_t *tiles, uint64_t T, uint32_t ts, uint32_t dim, uint8_t *out);
int emo_analyse_dev(emo_ctx *ctx, const uint8_t *tiles_dev, uint64_t T, uint32_t ts, uint32_t dim, uint8_t *out_dev);
/* One pass over the tiles producing both the 1to1 ([T,1,3]) and 4to1 ([T,4,3]) analyses
 * (what `-f` needs when both .emosaic_1to1 and .emosaic_4to1 are rebuilt). ts must be even. */
int emo_analyse_fused(emo_ctx *ctx, const uint8_t *tiles, uint64_t T, uint32_t ts, uint8_t *out1, uint8_t *out4);
int emo_analyse_fused_dev(emo_ctx *ctx, const uint8_t *tiles_dev, uint64_t T, uint32_t ts, uint8_t *out1_dev,
                          uint8_t *out4_dev);

/* ---- (2) library / search set -----------------------------------------------------------
 * Replaces TileSet::build_kiddo() (src/mosaic/tiles/tileset.rs:178-190) + Tile::coords()
 * (tiles/tile.rs:106-119) + flipped_coords() (tiles/utils.rs:18-43): every tile enters the
 * candidate set twice, (coords,+idx) then (mirror(coords),-idx), idx = position+1.
 * colors [T,N,3] (N = dim*dim analysis cells), tile_px [T,ts,ts,3] (may be NULL if only
 * emo_match is used).  The library stays resident on the GPU until replaced.
 * The reference's i16 limit (T <= 32767, tileset.rs:182) does not apply: T < 2^30. */
int emo_set_library(emo_ctx *ctx, const uint8_t *colors, const uint8_t *tile_px, uint32_t T, uint32_t N, uint32_t ts);
int emo_set_library_dev(emo_ctx *ctx, const uint8_t *colors_dev, const uint8_t *tile_px_dev, uint32_t T, uint32_t N,
                        uint32_t ts);
/* Shape of the resident library (all 0 when none is set; ts == 0 without tile pixels) — what a rank that received its
 * library from emo_comm_set_library needs to size its buffers.  Any output may be NULL. */
int emo_library_info(emo_ctx *ctx, uint32_t *T, uint32_t *N, uint32_t *ts);

/* ---- (2b) search index (1to1) ---------------------------------------------------------------
 * The GPU analogue of the KD-tree that TileSet::build_kiddo() (tiles/tileset.rs:178-190) returns and
 * render_nto1 (rendering.rs:136) builds once per render.  For N == 1 the query space is the 2^24 RGB
 * colours: emo_build_index computes, for every colour, the tile the scan would pick (exact city-block
 * distance transform over the colour cube, 64 MiB, same distance and tie-break), after which
 * emo_match is one table load per pixel.  Results are bit-identical with and without the index.
 * emo_build_index: builds (or rebuilds) the index of the resident library; EMO_ERR_STATE without a
 *   library, EMO_ERR_UNSUPPORTED when N != 1 or T > 2^22 (those libraries are always scanned).
 * emo_set_match_mode: EMO_MATCH_AUTO (default: use the index if it exists, build it on the first
 *   1to1 match whose scan would cost more than the build, i.e. blocks x tiles >= 2^31),
 *   EMO_MATCH_SCAN (always the brute-force scan kernel), EMO_MATCH_INDEX (always the index when the
 *   library supports one).  emo_set_library drops the index.
 * The index has two forms with identical answers: the 64 MiB table (key = distance and tile per colour) and a compact
 * 32 MiB one (u16 winner slot per colour, distance recomputed from the winner's colour) that stays L2-resident next to
 * the output stream; the library picks per launch.  EMO_MATCH_INDEX_WIDE / EMO_MATCH_INDEX_COMPACT are EMO_MATCH_INDEX
 * with that choice forced (tuning and tests; COMPACT falls back to WIDE for libraries with more than 65 536 distinct
 * colours). */
enum { EMO_MATCH_AUTO = 0, EMO_MATCH_SCAN = 1, EMO_MATCH_INDEX = 2, EMO_MATCH_INDEX_WIDE = 3, EMO_MATCH_INDEX_COMPACT = 4 };
int emo_build_index(emo_ctx *ctx);
int emo_set_match_mode(emo_ctx *ctx, int mode);

/* ---- (3) nearest-colour match -----------------------------------------------------------
 * Replaces the get_tile closure of render_nto1 (src/mosaic/rendering.rs:158-221):
 * get_img_colors (analysis.rs:23-36) -> coords -> KdTree::nearest_one::<Manhattan>
 * (rendering.rs:187-195) for every dim x dim block of src [H,W,3].
 * item/dist are [H/dim, W/dim] row-major; dist = L1 distance (sum |q-c| over 3N bytes);
 * ties: smallest idx, unflipped before flipped (kiddo's leaf-scan order).
 * W and H must be multiples of dim (main.rs:603-611). A source stripe (a contiguous range
 * of block rows) is just a smaller image, which is how multi-GPU runs shard the work. */
int emo_match(emo_ctx *ctx, const uint8_t *src, uint32_t W, uint32_t H, int32_t *item, uint32_t *dist);
int emo_match_dev(emo_ctx *ctx, const uint8_t *src_dev, uint32_t W, uint32_t H, int32_t *item_dev,
                  uint32_t *dist_dev);

/* ---- (3b) ranked candidates for the no-repeat renderer --------------------------------------
 * Replaces the Scoring phase of render_nto1_no_repeat (src/mosaic/rendering.rs:307-321,
 * nearest_n::<Manhattan>(coords, 100000) per block, and the refill compute_nearest(n, 10) at
 * :384-386).  For every dim x dim block of src: the candidates at positions [first, first + k) of
 * the block's candidate list sorted by (L1 distance, insertion rank) — rank 2t = tile t, 2t + 1 =
 * its mirror (tileset.rs:178-190; the order among equal distances is this library's canonical
 * one).  item_out/dist_out are [H/dim * W/dim][k], blocks row-major (by * bw + bx);
 * positions past the end of the list hold item 0 / dist 0xFFFFFFFF.  1 <= k <= 1024.
 * exclude (NULL or [T] bytes): 1 = tile retired; its candidates (both orientations) are left out of
 * every list, as the reference removes a placed tile from the tree (rendering.rs:366-380) before the
 * next refill.
 * N = 1, 4, 9, 16 (--mode 1..4); EMO_ERR_UNSUPPORTED otherwise.  For N == 1 the mirrored twin of a
 * tile (same vector, directly behind it in every list) is omitted: using a tile retires both
 * orientations (rendering.rs:357-358).  The greedy assignment that consumes the lists
 * (rendering.rs:341-392) is sequential and lives in the host layers (render_nto1_no_repeat). */
int emo_topk(emo_ctx *ctx, const uint8_t *src, uint32_t W, uint32_t H, uint32_t first, uint32_t k,
             const uint8_t *exclude, int32_t *item_out, uint32_t *dist_out);
int emo_topk_dev(emo_ctx *ctx, const uint8_t *src_dev, uint32_t W, uint32_t H, uint32_t first, uint32_t k,
                 const uint8_t *exclude_dev, int32_t *item_out_dev, uint32_t *dist_out_dev);

/* ---- (3c) no-repeat assignment ---------------------------------------------------------------
 * Replaces the scoring + greedy loop of render_nto1_no_repeat (src/mosaic/rendering.rs:262-392) in one call: every block
 * gets the nearest tile that no nearer block has taken; a tile is used once, in either orientation (:357-358).  Blocks are
 * served in increasing (distance of their best remaining candidate, block number n = bx * vtiles + by) order — the
 * reference's sorted vector (:323-326, :380-391) with this library's canonical order among equal distances.
 * The ranked lists come from the GPU (the kernel behind emo_topk): a deep first page for every block, and one batched launch
 * per refill epoch for the blocks that run dry (filtered by the tiles already placed, like the reference's pruned tree,
 * :384-386); the sequential merge runs on the host inside this call.
 * src [H,W,3] host; item / dist [H/dim, W/dim] host; item 0 = the block ran out of candidates and stays black (:347-351).
 * page: candidates per block of the first page, 0 = as deep as a 512 MB host budget allows (at most 1024).
 * counters (NULL or [4]): refill launches, blocks refilled, heap pops, tiles placed.
 * EMO_ERR_ARG when there are more blocks than 2T candidates (:292-298); N = 1, 4, 9, 16 like emo_topk. */
int emo_no_repeat(emo_ctx *ctx, const uint8_t *src, uint32_t W, uint32_t H, uint32_t page, int32_t *item, uint32_t *dist,
                  uint64_t *counters);

/* ---- (4) compose (+tint) -----------------------------------------------------------------
 * Replaces render() (src/mosaic/rendering.rs:51-101) + TileSet::get_image()
 * (tiles/tileset.rs:146-161) and, when out_channels == 4, the tint block
 * (src/main.rs:447-478: alpha overlay of the nearest-resized source, image 0.25.2
 * Rgba::blend in f32, bit-exact).
 * item [H/dim,W/dim] -> out [H/dim*ts, W/dim*ts, out_channels].
 *   out_channels == 3: RGB mosaic, src/tint_alpha ignored (src may be NULL)
 *   out_channels == 4: RGBA = blend(mosaic, src upscaled, alpha = tint_alpha); alpha byte as
 *                      the reference computes it. tint_alpha = (255*opacity) as u8. */
int emo_compose(emo_ctx *ctx, const int32_t *item, const uint8_t *src, uint32_t W, uint32_t H,
                uint32_t out_channels, uint8_t tint_alpha, uint8_t *out);
int emo_compose_dev(emo_ctx *ctx, const int32_t *item_dev, const uint8_t *src_dev, uint32_t W, uint32_t H,
                    uint32_t out_channels, uint8_t tint_alpha, uint8_t *out_dev);

/* Tint with an overlay image of a different size than the matched source: src/main.rs:447-478 overlays the ORIGINAL
 * image, which n_to_1 (main.rs:567-595) may have resized before matching (--downsample > 1, or dimensions that are not
 * multiples of dim).  overlay [oh,ow,3]; (W, H) are the dimensions of the matched source that produced `item`;
 * out [H/dim*ts, W/dim*ts, 4].  Sampling: floor((X + 0.5) * ow / OW) in f32, image 0.25.2 resize(Nearest). */
int emo_compose_overlay(emo_ctx *ctx, const int32_t *item, uint32_t W, uint32_t H, const uint8_t *overlay, uint32_t ow,
                        uint32_t oh, uint8_t tint_alpha, uint8_t *out);
int emo_compose_overlay_dev(emo_ctx *ctx, const int32_t *item_dev, uint32_t W, uint32_t H, const uint8_t *overlay_dev,
                            uint32_t ow, uint32_t oh, uint8_t tint_alpha, uint8_t *out_dev);

/* ---- (5) render statistics from the maps ----------------------------------------------------------
 * Replaces the accumulation of RenderStats — one push_tile per block under a Mutex (src/mosaic/rendering.rs:211-214,
 * stats.rs:56-64) — and the reductions summarise() / render() make over it (stats.rs:87-139, :169-175), as one pass over the
 * item / dist maps a match produced: sums[0] = blocks that carry a tile (item != 0), sums[1] = the sum of their distances,
 * sums[2] = their largest distance; usage[t] (NULL or [T]) = how many blocks placed tile t + 1, in either orientation.
 * The two ordered lists of the summary (ten most used, ten worst) are left to the host layers.
 * emo_stats_dev: device pointers (e.g. the maps emo_match_dev / emo_mosaic_dev just wrote), asynchronous; sums_dev is
 * three 8-byte aligned uint64.  emo_stats: host pointers, synchronous.  EMO_ERR_ARG for an id beyond T. */
int emo_stats_dev(emo_ctx *ctx, const int32_t *item_dev, const uint32_t *dist_dev, uint64_t Q, uint32_t T, uint64_t *sums_dev,
                  uint32_t *usage_dev);
int emo_stats(emo_ctx *ctx, const int32_t *item, const uint32_t *dist, uint64_t Q, uint32_t T, uint64_t sums[3], uint32_t *usage);

/* ---- whole path ------------------------------------------------------------------------------
 * render_nto1 (+ tint) in one call (rendering.rs:124-239 + main.rs:447-478).
 * emo_mosaic: host buffers; H2D(src) -> match -> compose -> D2H(out) with the copies pipelined
 *   against the kernels in block-row chunks. item/dist may be NULL.
 * emo_mosaic_dev: device buffers, asynchronous on the ctx stream: emo_match_dev followed by
 *   emo_compose_dev in one call (item_dev and dist_dev are required). */
int emo_mosaic(emo_ctx *ctx, const uint8_t *src, uint32_t W, uint32_t H, uint32_t out_channels,
               uint8_t tint_alpha, int32_t *item, uint32_t *dist, uint8_t *out);
int emo_mosaic_dev(emo_ctx *ctx, const uint8_t *src_dev, uint32_t W, uint32_t H, uint32_t out_channels,
                   uint8_t tint_alpha, int32_t *item_dev, uint32_t *dist_dev, uint8_t *out_dev);
/* The host-pointer calls stage through device buffers the ctx grows on demand (a synchronise + free + allocate the first
 * time a larger image arrives).  emo_reserve sizes them ahead of time for sources up to W x H with the resident library
 * (emo_match, emo_compose, emo_mosaic with out_channels 3 or 4), so that no call on the timed path allocates. */
int emo_reserve(emo_ctx *ctx, uint32_t W, uint32_t H, uint32_t out_channels);

/* ---- multi-GPU -----------------------------------------------------------------------------------
 * The path shards without an exchange step.  The units are the reference's own parallel tasks: one block row of
 * the source per task in render() ((0..H).into_par_iter().step_by(step), src/mosaic/rendering.rs:68-89, strips
 * merged at :91-99) and one tile per task in the analysis build (src/main.rs:760-794).  GPU r of n takes the
 * contiguous range emo_stripe_bounds(units, n, r); every GPU holds the whole library, replicated once by an NCCL
 * broadcast over NVLink; stripes of the output are disjoint row ranges of ONE host image, and each GPU copies its
 * stripe straight to its row offset — that copy is the host-side concatenation, there is no second one.
 * No collective runs inside the match / compose loop.
 *
 * NCCL is loaded at run time (dlopen: $EMO_NCCL_LIB, else the libnccl.so.2 already in the process, else the
 * system's) the first time a call needs it, so single-GPU hosts do not need NCCL installed; EMO_ERR_NCCL if absent. */

/* [start, stop) of `units` for part `rank` of `world`: contiguous, sizes differ by at most one. */
void emo_stripe_bounds(uint64_t units, int world, int rank, uint64_t *start, uint64_t *stop);

/* Pinned host memory for images the multi-GPU calls copy to and from.  emo_host_register pins memory the caller
 * already owns (e.g. a Rust Vec<u8> behind an RgbImage) for the duration of a render. */
int emo_host_register(void *p, size_t bytes);
int emo_host_unregister(void *p);

/* (a) one process per GPU (torchrun, MPI, ...).  Rank 0 calls emo_comm_unique_id and hands the 128 bytes to every
 * rank by any means (a file, torchrun's store, MPI_Bcast); every rank then calls emo_comm_init_rank on its ctx
 * (collective: ncclCommInitRank).  After that:
 *   emo_comm_set_library[_dev]: emo_set_library on every rank from the root's data.  On the root the arguments are
 *     those of emo_set_library[_dev]; on the other ranks colors / tile_px / T / N / ts are ignored (pass NULL / 0):
 *     the root's sizes travel in a header, the colours and tile pixels in one ncclBroadcast each.
 *   emo_comm_broadcast_dev: ncclBroadcast of a device buffer (e.g. the source image), in place.
 *   emo_comm_allgather_analysis_dev: the sharded analysis build — rank r analysed the tiles of
 *     emo_stripe_bounds(T, world, r) into local_dev [(stop - start) * bytes_per_tile]; afterwards all_dev
 *     [T * bytes_per_tile] holds every rank's results in tile order on every rank (one ncclAllGather).
 * The *_dev collectives are asynchronous on the ctx stream like every *_dev call. */
#define EMO_COMM_ID_BYTES 128
int emo_comm_unique_id(void *id_out /* [EMO_COMM_ID_BYTES] */);
int emo_comm_init_rank(emo_ctx *ctx, const void *id, int rank, int world);
int emo_comm_info(emo_ctx *ctx, int *rank, int *world, int *nccl_version);
int emo_comm_set_library(emo_ctx *ctx, const uint8_t *colors, const uint8_t *tile_px, uint32_t T, uint32_t N, uint32_t ts,
                         int root);
int emo_comm_set_library_dev(emo_ctx *ctx, const uint8_t *colors_dev, const uint8_t *tile_px_dev, uint32_t T, uint32_t N,
                             uint32_t ts, int root);
int emo_comm_broadcast_dev(emo_ctx *ctx, void *buf_dev, size_t bytes, int root);
int emo_comm_allgather_analysis_dev(emo_ctx *ctx, const uint8_t *local_dev, uint64_t T, uint32_t bytes_per_tile,
                                    uint8_t *all_dev);

/* (b) one process, several GPUs: an emo_group owns one ctx per device and (for n > 1) their communicators
 * (ncclCommInitAll).  Group calls take HOST pointers, are synchronous on return and run one worker thread per GPU.
 *   emo_group_set_library: H2D on the first device, ncclBroadcast to the others, search set built on every GPU.
 *   emo_group_analyse[_fused]: tiles sharded by emo_stripe_bounds(T, n, r); every GPU writes its range of `out`.
 *   emo_group_mosaic: block rows sharded by emo_stripe_bounds(H / dim, n, r); GPU r uploads its source stripe,
 *     matches, composes and copies its output stripe (and its rows of item / dist, which may be NULL) to its row
 *     offset in `out`.  Results are bit-identical to emo_mosaic on one GPU.
 * emo_group_ctx gives access to one member for the single-GPU calls (e.g. emo_set_match_mode, emo_launch_count). */
typedef struct emo_group emo_group;
int emo_group_create(const int *devices, int n, emo_group **out); /* devices == NULL: ordinals 0..n-1 */
void emo_group_destroy(emo_group *g);
int emo_group_size(const emo_group *g);
emo_ctx *emo_group_ctx(emo_group *g, int i);
int emo_group_set_library(emo_group *g, const uint8_t *colors, const uint8_t *tile_px, uint32_t T, uint32_t N, uint32_t ts);
int emo_group_analyse(emo_group *g, const uint8_t *tiles, uint64_t T, uint32_t ts, uint32_t dim, uint8_t *out);
int emo_group_analyse_fused(emo_group *g, const uint8_t *tiles, uint64_t T, uint32_t ts, uint8_t *out1, uint8_t *out4);
int emo_group_mosaic(emo_group *g, const uint8_t *src, uint32_t W, uint32_t H, uint32_t out_channels, uint8_t tint_alpha,
                     int32_t *item, uint32_t *dist, uint8_t *out);

#ifdef __cplusplus
}
#endif
#endif /* EMOSAIC_CUDA_H */
