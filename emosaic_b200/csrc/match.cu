// match.cu — search-set build and brute-force L1 nearest-colour matching.
//
// Reference: TileSet::build_kiddo() src/mosaic/tiles/tileset.rs:178-190 (every tile enters the
// KD-tree twice: (coords,+idx) then (mirror(coords),-idx)), Tile::coords() tiles/tile.rs:106-119,
// flipped_coords() tiles/utils.rs:18-43, get_img_colors() analysis.rs:23-36 and the query
// kdtree.nearest_one::<Manhattan>() in render_nto1's closure, rendering.rs:158-221.
//
// Semantics reproduced exactly: dist = sum over the 3N bytes of |q - c| (u32); winner = minimum
// distance, ties broken by insertion rank (rank 2t = tile t unflipped, 2t+1 = tile t mirrored),
// i.e. smallest idx first and unflipped before flipped — kiddo's leaf scan replaces the best only
// on strict `<`.  item = +(t+1) / -(t+1).
//
// Why CUDA cores: L1 distance is not bilinear, so there is no exact tensor-core contraction; the
// only exact embedding (thermometer codes, |a-b| = popc(ta^tb)) needs K = 765*N per pair and an
// epilogue that still does one min per pair on the CUDA cores, which costs more than this kernel's
// whole inner loop.  The inner loop is byte-SIMD instead: one VABSDIFF4.U8.ACC per 4 bytes of
// vector per pair, half an IMAD (FMA pipe) and a quarter of a VIMNMX3.U16x2 per pair.
//
// Data layout in HBM: candidates packed [Lpad][WORDS] u32 (WORDS = ceil(3N/4), bytes of the
// 3N-vector little-endian, zero padded), padded to a multiple of the stage size with copies of
// the last real candidate (a duplicate of an earlier candidate can never win under strict `<`); the last
// stage is scanned up to the last whole window that holds a real candidate, the rest of the padding is only there to be loadable.
// A CTA keeps NT*R queries in registers and streams the whole candidate array through a
// 4-stage shared-memory ring filled by TMA bulk copies issued from a dedicated producer warp;
// every candidate word is read with a warp-broadcast LDS.128 and reused for R queries.
// The running minimum is distance-only (two 16-bit minima packed per register); the argmin is
// recovered lazily: after each window of 128 candidates the thread checks whether any packed
// minimum changed and, if the overall minimum dropped, remembers the window.  An epilogue re-reads
// that one window from L2 and takes the first candidate that reaches the minimum (strict `<` on
// windows + first-in-window = smallest rank).  ALU pipe: 1.32 instructions per pair measured.
#include <stdlib.h>

#include "common.cuh"

static constexpr int MATCH_STAGES = 4;
static constexpr int MATCH_WIN = 128;  // candidates between two argmin checks
static constexpr int MATCH_UNROLL = 8;


// ---------------------------------------------------------------------------------------
// library build: packed candidates and the (tile, mirrored tile) pixel store
// ---------------------------------------------------------------------------------------
__global__ void build_candidates_kernel(const uint8_t *__restrict__ colors, uint32_t T, uint32_t N, uint32_t dim,
                                        uint32_t words, uint32_t L, uint32_t Lpad, uint32_t *__restrict__ cand) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= Lpad) return;
    const uint32_t cc = c < L ? c : L - 1;
    const uint32_t t = (N == 1) ? cc : (cc >> 1);
    const bool flipped = (N != 1) && (cc & 1);
    const uint8_t *v = colors + (size_t)t * N * 3;
    for (uint32_t w = 0; w < words; w++) {
        uint32_t packed = 0;
        for (uint32_t k = 0; k < 4; k++) {
            const uint32_t b = w * 4 + k;  // byte index in the 3N vector
            if (b < 3 * N) {
                uint32_t cell = b / 3, ch = b % 3;
                if (flipped) {  // utils.rs:18-43: swap cell column j <-> dim-1-j inside each cell row
                    const uint32_t row = cell / dim, col = cell % dim;
                    cell = row * dim + (dim - 1 - col);
                }
                packed |= (uint32_t)v[cell * 3 + ch] << (8 * k);
            }
        }
        cand[(size_t)c * words + w] = packed;
    }
}

// lib_px[2t] = tile t, lib_px[2t+1] = tile t mirrored horizontally (tileset.rs:156-157 flip_horizontal),
// so that compose is a pure gather.
__global__ void build_pixels_kernel(const uint8_t *__restrict__ px, uint32_t T, uint32_t ts, uint8_t *__restrict__ lib) {
    const uint64_t total = (uint64_t)T * ts * ts;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t t = i / ((uint64_t)ts * ts);
        const uint32_t rem = (uint32_t)(i % ((uint64_t)ts * ts));
        const uint32_t r = rem / ts, c = rem % ts;
        const uint8_t *s = px + i * 3;
        const uint8_t a = s[0], b = s[1], d = s[2];
        uint8_t *o0 = lib + ((size_t)(2 * t) * ts * ts + (size_t)r * ts + c) * 3;
        uint8_t *o1 = lib + ((size_t)(2 * t + 1) * ts * ts + (size_t)r * ts + (ts - 1 - c)) * 3;
        o0[0] = a; o0[1] = b; o0[2] = d;
        o1[0] = a; o1[1] = b; o1[2] = d;
    }
}

int emo_launch_tile_candidates(emo_ctx *ctx);

int emo_launch_build_library(emo_ctx *ctx, const uint8_t *colors, const uint8_t *tile_px) {
    const uint32_t Lpad = ctx->n_chunks * ctx->chunk;
    build_candidates_kernel<<<(Lpad + 255) / 256, 256, 0, ctx->stream>>>(colors, ctx->T, ctx->N, ctx->dim, ctx->words, ctx->L,
                                                                        Lpad, ctx->cand);
    EMO_LAUNCH_CHECK(ctx);
    if (ctx->wide) {
        int rc = emo_launch_tile_candidates(ctx);
        if (rc) return rc;
    }
    if (tile_px) {
        uint64_t total = (uint64_t)ctx->T * ctx->ts * ctx->ts;
        uint64_t blocks = (total + 255) / 256, cap = (uint64_t)ctx->sm_count * 16;
        build_pixels_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, ctx->stream>>>(tile_px, ctx->T, ctx->ts, ctx->lib_px);
        EMO_LAUNCH_CHECK(ctx);
    }
    return EMO_OK;
}

// ---------------------------------------------------------------------------------------
// match kernel
// ---------------------------------------------------------------------------------------
struct MatchParams {
    const uint32_t *cand;   // [n_chunks*chunk][WORDS]
    uint32_t chunk;         // candidates per stage (multiple of MATCH_WIN)
    uint32_t n_chunks;      // total stages in the candidate array
    uint32_t last_len;      // candidates scanned in the last stage: what is left of L, rounded up to whole windows
    uint32_t publish_all;   // 1: whole-range CTAs also leave the search inside the winning window to match_finalize_window_kernel
    uint32_t chunks_per_split;  // stages per split CTA
    uint32_t split_ctas;        // the first split_ctas CTAs of the grid share query tiles: tile = split_tile0 + x / splits, part = x % splits
    uint32_t splits, split_tile0;
    const uint8_t *src;     // [H][W][3]
    uint32_t W, bw, Q, dim;
    uint32_t mirrored;      // 1: candidate c -> tile c>>1, flipped c&1 ; 0: candidate c -> tile c
    int32_t *item;
    uint32_t *dist;
    unsigned long long *keys;  // non-null: split mode, atomicMin of (dist<<32 | candidate)
};

template <int WORDS, int OFF, int NCW>
__device__ __forceinline__ uint32_t sad_vec(const uint32_t (&q)[WORDS], const uint32_t (&c)[NCW], uint32_t acc) {
    uint32_t d = acc;
#pragma unroll
    for (int w = 0; w < WORDS; w++) d = sad4(q[w], c[OFF + w], d);
    return d;
}

// resident CTAs per SM the register budget is planned for
template <int WORDS, int R, int NT>
constexpr int match_min_blocks() {
    return R >= 8 ? (NT >= 256 ? 2 : 4) : (WORDS > 4 ? 3 : (R >= 4 ? (NT <= 64 ? 8 : 4) : 6));
}

// WIN: candidates between two argmin checks (the stage size must be a multiple); MINB: resident CTAs per SM the registers
// are capped for (0: match_min_blocks); UNR: unroll factor of the 4-candidate loop
template <int WORDS, int R, int NT, int WIN = MATCH_WIN, int MINB = 0, int UNR = MATCH_UNROLL>
__global__ void __launch_bounds__(NT + 32, (MINB ? MINB : match_min_blocks<WORDS, R, NT>())) match_kernel(const MatchParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t stage_words = p.chunk * WORDS;
    uint32_t *ring = reinterpret_cast<uint32_t *>(smem);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)MATCH_STAGES * stage_words * 4);
    uint64_t *empty = full + MATCH_STAGES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int CONSUMER_WARPS = NT / 32;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < MATCH_STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], CONSUMER_WARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();

    // Grid = [split CTAs | whole-range CTAs].  A split CTA scans one part of the candidate range for a query tile it shares with
    // `splits - 1` others and publishes (distance, window); a whole-range CTA owns its tile and finishes it itself.  The split
    // CTAs come first so that they run next to the long ones instead of after them.
    const bool split = blockIdx.x < p.split_ctas;
    const uint32_t tile = split ? p.split_tile0 + blockIdx.x / p.splits : blockIdx.x - p.split_ctas;
    const uint32_t c0 = split ? (blockIdx.x % p.splits) * p.chunks_per_split : 0u;
    const uint32_t c1 = split ? min(c0 + p.chunks_per_split, p.n_chunks) : p.n_chunks;

    if (warp == CONSUMER_WARPS) {
        // ---- producer warp: one lane streams candidate stages through the ring with TMA bulk copies
        if (lane == 0) {
            for (uint32_t c = c0, n = 0; c < c1; c++, n++) {
                const int s = n % MATCH_STAGES;
                const uint32_t bytes = (c + 1 == p.n_chunks ? p.last_len : p.chunk) * (WORDS * 4);
                if (n >= MATCH_STAGES) mbar_wait_relaxed(&empty[s], ((n / MATCH_STAGES) - 1) & 1);
                mbar_arrive_expect_tx(&full[s], bytes);
                bulk_g2s(ring + (size_t)s * stage_words, p.cand + (size_t)c * stage_words, bytes, &full[s]);
            }
        }
        return;
    }

    // ---- consumers: R queries per thread, packed like the candidates
    // best2[r] holds TWO running minima as 16-bit lanes: high lane = candidates at even positions,
    // low lane = odd positions.  A pair of distances is produced already packed:
    //   v = sad(q, c_odd, sad(q, c_even, 0) << 16)      (VABSDIFF4.ACC, IMAD.U32 x 0x10000, VABSDIFF4.ACC)
    // and one VIMNMX3.U16x2 folds four candidates into best2.  ALU pipe: 1.25 instr/pair, FMA pipe: 0.5.
    // Distances are < 65536 because 255 * 3N <= 12240 for the supported N.
    uint32_t q[R][WORDS], best2[R], seen2[R], bestd[R], idx[R];
    const uint32_t qbase = tile * (uint32_t)(NT * R);
#pragma unroll
    for (int r = 0; r < R; r++) {
        uint32_t qi = qbase + r * NT + tid;
        if (qi >= p.Q) qi = p.Q - 1;
        const uint32_t by = qi / p.bw, bx = qi % p.bw;
#pragma unroll
        for (int w = 0; w < WORDS; w++) q[r][w] = 0;
        // analysis.rs:23-36: cell i of the block is source pixel (x + i % dim, y + i / dim);
        // byte b of the query vector is channel b % 3 of cell b / 3 (all register indices static)
        const uint32_t D = 3 * p.dim * p.dim;
#pragma unroll
        for (int b = 0; b < 4 * WORDS; b++) {
            if ((uint32_t)b < D) {
                const uint32_t cell = b / 3, ch = b % 3;
                const uint32_t cy = cell / p.dim, cx = cell - cy * p.dim;
                const uint32_t v = p.src[((size_t)(by * p.dim + cy) * p.W + (bx * p.dim + cx)) * 3 + ch];
                q[r][b >> 2] |= v << (8 * (b & 3));
            }
        }
        best2[r] = 0xffffffffu;
        seen2[r] = 0xffffffffu;
        bestd[r] = 0xffffffffu;
        idx[r] = 0;
    }

    for (uint32_t c = c0, n = 0; c < c1; c++, n++) {
        const int s = n % MATCH_STAGES;
        mbar_wait(&full[s], (n / MATCH_STAGES) & 1);
        const uint32_t *st = ring + (size_t)s * stage_words;
        const uint32_t cand_base = c * p.chunk;
        const uint32_t len = c + 1 == p.n_chunks ? p.last_len : p.chunk;  // the padding behind the last window is not scanned
        for (uint32_t w0 = 0; w0 < len; w0 += WIN) {
            const uint4 *win = reinterpret_cast<const uint4 *>(st + (size_t)w0 * WORDS);
#pragma unroll(UNR)
            for (int j4 = 0; j4 < WIN / 4; j4++) {
                // 4 candidates = 4*WORDS words = WORDS x LDS.128 (same address in every lane: broadcast)
                uint32_t cw[4 * WORDS];
#pragma unroll
                for (int v = 0; v < WORDS; v++) {
                    const uint4 t4 = win[j4 * WORDS + v];
                    cw[4 * v + 0] = t4.x; cw[4 * v + 1] = t4.y; cw[4 * v + 2] = t4.z; cw[4 * v + 3] = t4.w;
                }
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const uint32_t e0 = sad_vec<WORDS, 0 * WORDS>(q[r], cw, 0u);
                    const uint32_t e2 = sad_vec<WORDS, 2 * WORDS>(q[r], cw, 0u);
                    const uint32_t v01 = sad_vec<WORDS, 1 * WORDS>(q[r], cw, e0 << 16);
                    const uint32_t v23 = sad_vec<WORDS, 3 * WORDS>(q[r], cw, e2 << 16);
                    best2[r] = __vimin3_u16x2(best2[r], v01, v23);  // VIMNMX3.U16x2
                }
            }
            // Lazy argmin: remember only WHICH window last lowered the overall minimum (no rescan here).
            bool changed = false;
#pragma unroll
            for (int r = 0; r < R; r++) changed |= best2[r] != seen2[r];
            if (changed) {
                const uint32_t wpos = cand_base + w0;
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const uint32_t m = min(best2[r] >> 16, best2[r] & 0xffffu);
                    seen2[r] = best2[r];
                    if (m < bestd[r]) { bestd[r] = m; idx[r] = wpos; }  // strict `<`: an equal later window never wins
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }

    // Split mode (several CTAs share a query tile): publish (distance, window) and leave the search inside the window to
    // match_finalize_window_kernel — a CTA that scans a short candidate range would otherwise spend most of its life in the
    // latency-bound rescan below.  Windows are disjoint index ranges, so the 64-bit minimum is still (distance, smallest index).
    if (split) {
#pragma unroll
        for (int r = 0; r < R; r++) {
            const uint32_t qi = qbase + r * NT + tid;
            if (qi < p.Q) atomicMin(&p.keys[qi], ((unsigned long long)bestd[r] << 32) | idx[r]);
        }
        return;
    }
    // A launch of few waves ends with every resident CTA in the rescan below at the same time (16 dependent rounds of L2 loads
    // per query with nothing to overlap them); the finalize kernel does the same search coalesced, one warp per query.
    if (p.publish_all) {
#pragma unroll
        for (int r = 0; r < R; r++) {
            const uint32_t qi = qbase + r * NT + tid;
            if (qi < p.Q) p.keys[qi] = ((unsigned long long)bestd[r] << 32) | idx[r];
        }
        return;
    }
    // Epilogue: the winner is the FIRST candidate of the remembered window that reaches the minimum
    // (candidates are scanned in rank order, so this is the canonical smallest-rank tie-break).
    // The window is re-read from global memory (L2-resident) with batched 128-bit loads.
#pragma unroll 1
    for (int r = 0; r < R; r++) {
        // runtime-indexed copy of this query (R is small; the selects are cheap and happen once per query)
        uint32_t qr[WORDS], br = 0, ir = 0;
#pragma unroll
        for (int rr = 0; rr < R; rr++)
            if (rr == r) {
#pragma unroll
                for (int w = 0; w < WORDS; w++) qr[w] = q[rr][w];
                br = bestd[rr];
                ir = idx[rr];
            }
        const uint4 *wc = reinterpret_cast<const uint4 *>(p.cand + (size_t)ir * WORDS);
        uint32_t found = WIN - 1;
        constexpr int GB = WORDS <= 2 ? 4 : (WORDS <= 4 ? 2 : 1);  // groups of 4 candidates loaded per batch
        for (int g0 = WIN / 4 - GB; g0 >= 0; g0 -= GB) {
            uint4 buf[GB * WORDS];
#pragma unroll
            for (int i = 0; i < GB * WORDS; i++) buf[i] = __ldg(wc + g0 * WORDS + i);
#pragma unroll
            for (int g = GB - 1; g >= 0; g--) {
                uint32_t cw[4 * WORDS];
#pragma unroll
                for (int v = 0; v < WORDS; v++) {
                    cw[4 * v + 0] = buf[g * WORDS + v].x; cw[4 * v + 1] = buf[g * WORDS + v].y;
                    cw[4 * v + 2] = buf[g * WORDS + v].z; cw[4 * v + 3] = buf[g * WORDS + v].w;
                }
                if (sad_vec<WORDS, 3 * WORDS>(qr, cw, 0u) == br) found = (g0 + g) * 4 + 3;
                if (sad_vec<WORDS, 2 * WORDS>(qr, cw, 0u) == br) found = (g0 + g) * 4 + 2;
                if (sad_vec<WORDS, 1 * WORDS>(qr, cw, 0u) == br) found = (g0 + g) * 4 + 1;
                if (sad_vec<WORDS, 0 * WORDS>(qr, cw, 0u) == br) found = (g0 + g) * 4 + 0;
            }
        }
#pragma unroll
        for (int rr = 0; rr < R; rr++)
            if (rr == r) idx[rr] = ir + found;
    }

#pragma unroll
    for (int r = 0; r < R; r++) {
        const uint32_t qi = qbase + r * NT + tid;
        if (qi >= p.Q) continue;
        const uint32_t cnd = idx[r];
        const int32_t t1 = (int32_t)(p.mirrored ? (cnd >> 1) : cnd) + 1;
        p.item[qi] = (p.mirrored && (cnd & 1)) ? -t1 : t1;
        p.dist[qi] = bestd[r];
    }
}

__global__ void match_init_keys_kernel(unsigned long long *keys, uint32_t Q, uint32_t q_begin = 0) {
    const uint32_t i = q_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i < Q) keys[i] = ~0ull;
}

// keys[q] = distance << 32 | candidate (the wide kernel's splits merge exact candidates)
__global__ void match_finalize_kernel(const unsigned long long *__restrict__ keys, uint32_t Q, uint32_t mirrored,
                                      int32_t *__restrict__ item, uint32_t *__restrict__ dist) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Q) return;
    const unsigned long long k = keys[i];
    const uint32_t cnd = (uint32_t)k;
    const int32_t t1 = (int32_t)(mirrored ? (cnd >> 1) : cnd) + 1;
    item[i] = (mirrored && (cnd & 1)) ? -t1 : t1;
    dist[i] = (uint32_t)(k >> 32);
}

// Split mode, second half: keys[q] = distance << 32 | first candidate of the window that holds the winner.  Eight lanes per
// query: lane l of the group takes candidates 16l .. 16l + 15 of the window (WIN = 128) in four batches of 128-bit loads, the
// first candidate whose distance equals the minimum is the canonical winner (smallest rank).  Coalesced, no dependent chain.
// (One warp per query spent most of its instructions assembling the same query vector in all 32 lanes: 112 us for C2's
// 227 000 queries, half of them ALU-pipe cycles.)
template <int WORDS>
__global__ void __launch_bounds__(256) match_finalize_window_kernel(const MatchParams p, uint32_t q_begin) {
    static_assert(MATCH_WIN == 128, "eight lanes x 16 candidates");
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t sub = threadIdx.x & 7, grp = (threadIdx.x & 31) >> 3;
    uint32_t qi = q_begin + (t >> 3);
    const bool live = qi < p.Q;
    if (!live) qi = p.Q - 1;  // keeps the warp whole for the ballot
    const unsigned long long k = p.keys[qi];
    const uint32_t best = (uint32_t)(k >> 32), w0 = (uint32_t)k;
    uint32_t q[WORDS];
#pragma unroll
    for (int w = 0; w < WORDS; w++) q[w] = 0;
    const uint32_t by = qi / p.bw, bx = qi % p.bw, D = 3 * p.dim * p.dim;
#pragma unroll
    for (int b = 0; b < 4 * WORDS; b++) {
        if ((uint32_t)b < D) {
            const uint32_t cell = b / 3, ch = b % 3;
            const uint32_t cy = cell / p.dim, cx = cell - cy * p.dim;
            const uint32_t v = p.src[((size_t)(by * p.dim + cy) * p.W + (bx * p.dim + cx)) * 3 + ch];
            q[b >> 2] |= v << (8 * (b & 3));
        }
    }
    const uint4 *wc = reinterpret_cast<const uint4 *>(p.cand + ((size_t)w0 + 16 * sub) * WORDS);
    uint32_t found = 16;
#pragma unroll
    for (int g = 3; g >= 0; g--) {
        uint32_t cw[4 * WORDS];
#pragma unroll
        for (int v = 0; v < WORDS; v++) {
            const uint4 t4 = __ldg(wc + g * WORDS + v);
            cw[4 * v + 0] = t4.x; cw[4 * v + 1] = t4.y; cw[4 * v + 2] = t4.z; cw[4 * v + 3] = t4.w;
        }
        if (sad_vec<WORDS, 3 * WORDS>(q, cw, 0u) == best) found = 4 * g + 3;
        if (sad_vec<WORDS, 2 * WORDS>(q, cw, 0u) == best) found = 4 * g + 2;
        if (sad_vec<WORDS, 1 * WORDS>(q, cw, 0u) == best) found = 4 * g + 1;
        if (sad_vec<WORDS, 0 * WORDS>(q, cw, 0u) == best) found = 4 * g + 0;
    }
    const uint32_t hit = (__ballot_sync(0xFFFFFFFFu, found < 16) >> (8 * grp)) & 0xFFu;
    const uint32_t first = hit ? (uint32_t)__ffs((int)hit) - 1 : 0;  // hit != 0: the window was chosen because it reaches `best`
    const uint32_t pos = __shfl_sync(0xFFFFFFFFu, found, 8 * grp + first);
    if (sub == 0 && live) {
        const uint32_t cnd = w0 + 16 * first + (pos < 16 ? pos : 0);
        const int32_t t1 = (int32_t)(p.mirrored ? (cnd >> 1) : cnd) + 1;
        p.item[qi] = (p.mirrored && (cnd & 1)) ? -t1 : t1;
        p.dist[qi] = best;
    }
}

template <int WORDS, int R, int NT, int WIN = MATCH_WIN, int MINB = 0, int UNR = MATCH_UNROLL>
static int launch_match_t(emo_ctx *ctx, MatchParams &p, uint32_t Q) {
    auto kern = match_kernel<WORDS, R, NT, WIN, MINB, UNR>;
    static const size_t smem_pad = getenv("EMO_MATCH_SMEM_PAD") ? (size_t)atoi(getenv("EMO_MATCH_SMEM_PAD")) : 0;  // tuning: caps the resident CTAs
    const size_t smem = (size_t)MATCH_STAGES * p.chunk * WORDS * 4 + 2 * MATCH_STAGES * 8 + smem_pad;
    EMO_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t qtiles = (Q + NT * R - 1) / (NT * R);
    {
        const uint32_t left = ctx->L - (p.n_chunks - 1) * p.chunk;  // candidates of the last stage, in whole windows
        p.last_len = (left + WIN - 1) / WIN * WIN;
        if (p.last_len > p.chunk) p.last_len = p.chunk;
    }
    // Split the candidate range across gridDim.y when the query tiles cannot fill the resident CTA slots: a split CTA pays a
    // prologue (query gather, pipeline fill) but no epilogue any more (match_finalize_window_kernel), so short ranges are fine:
    // at least 4 stages of candidates per CTA.
    int occ = 1;
    EMO_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT + 32, smem));
    if (occ < 1) occ = 1;
    const uint32_t slots = (uint32_t)ctx->sm_count * (uint32_t)occ;
    // Which query tiles share their candidate range between several CTAs:
    //  - fewer tiles than resident slots: all of them, the number of parts from the cost model below (C2's row stripes,
    //    tools/sweep_match3.py: 64 block rows 204 us unsplit, 162 us in 8 parts; splitting a grid that already fills the slots costs 8 %);
    //  - a small ragged last wave (rem = tiles % slots <= slots / 8): only those rem tiles, slots / rem parts each, so that the
    //    remainder fills the machine next to the whole-range CTAs instead of running one CTA per SM after them.  Without it the
    //    time is a staircase — 888 tiles 954 us, anything from 889 to 1024 tiles 1105 us on C2's library — with it 896 tiles take
    //    978 us, 912: 1008, 960: 1060, 992: 1090; from rem ~ 120 on the lone CTAs are as cheap (1024 tiles: 1139 vs 1105 us).
    // A part is at least 4 stages long.
    uint32_t splits = 1, split_tiles = 0;
    const uint32_t max_parts = p.n_chunks / 4;
    static const bool tail_on = !(getenv("EMO_MATCH_TAIL") && atoi(getenv("EMO_MATCH_TAIL")) == 0);            // tuning / A-B switch
    if (qtiles < slots) {
        // Parts per tile from a cost model fitted to tools/sweep_match3.py on C2's stripes (EMO_MATCH_SPLITS = 2 ... 10 on 32 - 256 block
        // rows): an SM works through its ceil(CTAs / SMs) CTAs at a constant rate once it holds three or more, a CTA costs its
        // stages plus ~0.3 of a stage for the prologue and the publish, so time ~ ceil(tiles * S / SMs) * (ceil(stages / S) + 0.3).
        // Even parts matter more than their number: 40 stages in 8 parts beat 6 parts (7,7,7,7,7,5) by 14 % on 64 rows.
        uint64_t best_cost = ~0ull;
        const uint32_t hi = max_parts ? max_parts : 1;
        for (uint32_t s = 1; s <= hi; s++) {
            const uint64_t ctas = (uint64_t)qtiles * s, per_sm = (ctas + ctx->sm_count - 1) / ctx->sm_count;
            uint64_t cost = per_sm * (10ull * ((p.n_chunks + s - 1) / s) + 3);
            if (ctas < 3ull * ctx->sm_count) cost = cost * 108 / 100;  // too few warps per SM to keep the ALU pipe busy
            if (cost < best_cost) { best_cost = cost; splits = s; }
        }
        static const uint32_t min_len = getenv("EMO_MATCH_MINLEN") ? (uint32_t)atoi(getenv("EMO_MATCH_MINLEN")) : 0u;  // tuning override
        if (min_len && splits > (p.n_chunks * p.chunk) / min_len) splits = (p.n_chunks * p.chunk) / min_len;
        split_tiles = qtiles;
    } else if (tail_on) {
        const uint32_t rem = qtiles % slots;
        if (rem && 8 * rem <= slots) {
            splits = slots / rem < max_parts ? slots / rem : max_parts;
            split_tiles = rem;
        }
    }
    if (const char *e = getenv("EMO_MATCH_SPLITS")) {  // tuning override: every tile in that many parts
        const int v = atoi(e);
        if (v >= 1) { splits = (uint32_t)v < p.n_chunks ? (uint32_t)v : p.n_chunks; split_tiles = qtiles; }
    }
    if (splits < 2) { splits = 1; split_tiles = 0; }
    p.chunks_per_split = (p.n_chunks + splits - 1) / splits;
    splits = (p.n_chunks + p.chunks_per_split - 1) / p.chunks_per_split;  // every part non-empty
    if (splits < 2) { splits = 1; split_tiles = 0; }
    p.splits = splits;
    p.split_tile0 = qtiles - split_tiles;
    p.split_ctas = split_tiles * splits;
    const uint32_t q_split = p.split_tile0 * (uint32_t)(NT * R);  // queries from here on are merged through the keys
    // EMO_MATCH_PUBLISH=1 (A-B switch, off by default): the whole-range CTAs publish (distance, window) as well and the finalize
    // kernel searches every window.  Measured on C2 (1024 CTAs on 888 slots, which end in their rescans together): the scan
    // itself 946 -> 922 us, but the finalize kernel costs more than the 24 us it saves.
    static const bool publish_env = getenv("EMO_MATCH_PUBLISH") && atoi(getenv("EMO_MATCH_PUBLISH")) > 0;
    p.publish_all = WIN == MATCH_WIN && p.split_tile0 && publish_env;
    const uint32_t q_begin = p.publish_all ? 0u : q_split;
    if (split_tiles || p.publish_all) {
        int rc = emo_ensure(ctx, (void **)&ctx->keys, &ctx->keys_cap, (size_t)Q * 8);
        if (rc) return rc;
        p.keys = ctx->keys;
    } else {
        p.keys = nullptr;
    }
    if (split_tiles) {
        match_init_keys_kernel<<<(Q - q_split + 255) / 256, 256, 0, ctx->stream>>>(ctx->keys, Q, q_split);
        EMO_LAUNCH_CHECK(ctx);
    }
    kern<<<p.split_ctas + p.split_tile0, NT + 32, smem, ctx->stream>>>(p);
    EMO_LAUNCH_CHECK(ctx);
    if (p.keys) {
        match_finalize_window_kernel<WORDS><<<(Q - q_begin + 31) / 32, 256, 0, ctx->stream>>>(p, q_begin);
        EMO_LAUNCH_CHECK(ctx);
    }
    return EMO_OK;
}

// ---------------------------------------------------------------------------------------
// wide vectors (N = 25 ... 16384, --mode 5 ... 128): tiled all-pairs kernel
// ---------------------------------------------------------------------------------------
// Query vectors do not fit in registers, so this is a shared-memory tiling fed by TMA.  Both operands are
// stored pre-tiled and word-major in HBM:
//     candidates [c-tile of 128][slice of 8 words][word][candidate]     (built once in emo_set_library)
//     queries    [q-tile of 64 ][slice of 8 words][word][query]         (pack_queries_kernel, per call)
// so one slice of a tile is a contiguous 4 KB / 2 KB block: a producer warp streams (query slice, candidate
// slice) pairs through an 8-stage mbarrier ring with two cp.async.bulk copies per stage and nothing is
// transposed or swizzled on the SM.  Warp w owns queries 8w..8w+7 (two warp-uniform LDS.128 per word), lane l
// owns candidates 4l..4l+3 (one conflict-free LDS.128 per word): 3 LDS.128 feed 32 VABSDIFF4.ACC.
// Candidates are visited in increasing rank and compared with strict `<`, lanes merge with a lexicographic
// (dist, rank) shuffle minimum, splits with the 64-bit atomicMin: the canonical tie-break again.
static constexpr int WQ = 64, WC = 128;  // slice depth WK (words) = 8, 16 or 32: the largest that divides the padded vector

// dst layout [tile][slice][k][row_in_tile]; src bytes gathered per word
__global__ void pack_queries_kernel(const uint8_t *__restrict__ src, uint32_t W, uint32_t bw, uint32_t Q, uint32_t Qpad,
                                    uint32_t dim, uint32_t words, uint32_t WK, uint32_t *__restrict__ qvec) {
    const uint64_t total = (uint64_t)Qpad * words;
    const uint32_t D = 3 * dim * dim, n_slices = words / WK;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t row = (uint32_t)(i % WQ);
        const uint32_t k = (uint32_t)((i / WQ) % WK);
        const uint32_t sl = (uint32_t)((i / (WQ * WK)) % n_slices);
        const uint32_t tile = (uint32_t)(i / ((uint64_t)WQ * WK * n_slices));
        uint32_t qi = tile * WQ + row;
        if (qi >= Q) qi = Q - 1;
        const uint32_t w = sl * WK + k;
        const uint32_t by = qi / bw, bx = qi % bw;
        uint32_t packed = 0;
        for (uint32_t j = 0; j < 4; j++) {
            const uint32_t b = w * 4 + j;
            if (b < D) {
                const uint32_t cell = b / 3, ch = b % 3, cy = cell / dim, cx = cell % dim;  // analysis.rs:23-36
                packed |= (uint32_t)src[((size_t)(by * dim + cy) * W + (bx * dim + cx)) * 3 + ch] << (8 * j);
            }
        }
        qvec[i] = packed;
    }
}

// plain [Lpad][words] candidates -> [c-tile][slice][k][candidate]
__global__ void tile_candidates_kernel(const uint32_t *__restrict__ plain, uint32_t Lpad, uint32_t words, uint32_t WK,
                                       uint32_t *__restrict__ tiled) {
    const uint64_t total = (uint64_t)Lpad * words;
    const uint32_t n_slices = words / WK;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t row = (uint32_t)(i % WC);
        const uint32_t k = (uint32_t)((i / WC) % WK);
        const uint32_t sl = (uint32_t)((i / (WC * WK)) % n_slices);
        const uint32_t tile = (uint32_t)(i / ((uint64_t)WC * WK * n_slices));
        tiled[i] = plain[(size_t)(tile * WC + row) * words + sl * WK + k];
    }
}

template <int WK>
__global__ void __launch_bounds__(288, 2) match_wide_kernel(const uint32_t *__restrict__ qvec, const uint32_t *__restrict__ cand,
                                                            uint32_t n_slices, uint32_t n_ctiles, uint32_t tiles_per_split,
                                                            uint32_t Q, unsigned long long *__restrict__ keys) {
    constexpr int WIDE_STAGES = 64 / WK;                     // 8 / 4 / 2 stages of 6 / 12 / 24 KB
    constexpr int WIDE_STAGE_WORDS = WK * (WQ + WC);
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t *ring = reinterpret_cast<uint32_t *>(smem);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)WIDE_STAGES * WIDE_STAGE_WORDS * 4);
    uint64_t *empty = full + WIDE_STAGES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < WIDE_STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 8);
        }
        mbar_fence_init();
    }
    __syncthreads();
    const uint32_t t0 = blockIdx.y * tiles_per_split, t1 = min(t0 + tiles_per_split, n_ctiles);
    const uint32_t n_steps = (t1 - t0) * n_slices;
    if (warp == 8) {  // producer: one lane, two bulk copies per stage
        if (lane == 0) {
            const uint32_t *qbase = qvec + (size_t)blockIdx.x * n_slices * (WK * WQ);
            for (uint32_t n = 0; n < n_steps; n++) {
                const int s = n % WIDE_STAGES;
                const uint32_t t = t0 + n / n_slices, sl = n % n_slices;
                if (n >= WIDE_STAGES) mbar_wait_relaxed(&empty[s], ((n / WIDE_STAGES) - 1) & 1);
                mbar_arrive_expect_tx(&full[s], WIDE_STAGE_WORDS * 4);
                uint32_t *dst = ring + (size_t)s * WIDE_STAGE_WORDS;
                bulk_g2s(dst, qbase + (size_t)sl * (WK * WQ), WK * WQ * 4, &full[s]);
                bulk_g2s(dst + WK * WQ, cand + ((size_t)t * n_slices + sl) * (WK * WC), WK * WC * 4, &full[s]);
            }
        }
        return;
    }
    uint32_t acc[8][4], bestd[8], besti[8];
#pragma unroll
    for (int a = 0; a < 8; a++) {
        bestd[a] = 0xffffffffu;
        besti[a] = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b] = 0;
    }
    for (uint32_t n = 0; n < n_steps; n++) {
        const int s = n % WIDE_STAGES;
        mbar_wait(&full[s], (n / WIDE_STAGES) & 1);
        const uint32_t *sq = ring + (size_t)s * WIDE_STAGE_WORDS + warp * 8;
        const uint32_t *sc = ring + (size_t)s * WIDE_STAGE_WORDS + WK * WQ + lane * 4;
#pragma unroll
        for (int k = 0; k < WK; k++) {
            const uint4 a0 = *reinterpret_cast<const uint4 *>(sq + k * WQ);
            const uint4 a1 = *reinterpret_cast<const uint4 *>(sq + k * WQ + 4);
            const uint4 c4 = *reinterpret_cast<const uint4 *>(sc + k * WC);
            const uint32_t qa[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, cb[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
            for (int a = 0; a < 8; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) acc[a][b] = sad4(qa[a], cb[b], acc[a][b]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (n % n_slices == n_slices - 1) {  // tile finished: fold (increasing rank, strict `<`)
            const uint32_t cbase = (t0 + n / n_slices) * WC + lane * 4;
#pragma unroll
            for (int a = 0; a < 8; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    if (acc[a][b] < bestd[a]) { bestd[a] = acc[a][b]; besti[a] = cbase + b; }
                    acc[a][b] = 0;
                }
        }
    }
    // merge the 32 lanes (lexicographic (dist, rank) minimum), then across splits
#pragma unroll
    for (int a = 0; a < 8; a++) {
        unsigned long long key = ((unsigned long long)bestd[a] << 32) | besti[a];
#pragma unroll
        for (int m = 1; m < 32; m <<= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, key, m);
            key = o < key ? o : key;
        }
        const uint32_t qi = blockIdx.x * WQ + warp * 8 + a;
        if (lane == 0 && qi < Q && n_steps) atomicMin(&keys[qi], key);
    }
}

static uint32_t wide_slice_words(uint32_t words) { return words % 32 == 0 ? 32 : (words % 16 == 0 ? 16 : 8); }

int emo_launch_tile_candidates(emo_ctx *ctx) {  // called by emo_launch_build_library for wide libraries
    const uint32_t Lpad = ctx->n_chunks * ctx->chunk;
    int rc = emo_ensure(ctx, (void **)&ctx->qvec, &ctx->qvec_cap, (size_t)Lpad * ctx->words * 4);  // scratch for the plain layout
    if (rc) return rc;
    EMO_CK(cudaMemcpyAsync(ctx->qvec, ctx->cand, (size_t)Lpad * ctx->words * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    const uint64_t total = (uint64_t)Lpad * ctx->words, blocks = (total + 255) / 256, cap = (uint64_t)ctx->sm_count * 32;
    tile_candidates_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, ctx->stream>>>(ctx->qvec, Lpad, ctx->words,
                                                                                             wide_slice_words(ctx->words), ctx->cand);
    EMO_LAUNCH_CHECK(ctx);
    return EMO_OK;
}

static int launch_match_wide(emo_ctx *ctx, const uint8_t *src, uint32_t W, uint32_t H, int32_t *item, uint32_t *dist) {
    const uint32_t dim = ctx->dim, bw = W / dim, Q = bw * (H / dim), words = ctx->words;
    const uint32_t WK = wide_slice_words(words), n_slices = words / WK;
    const uint32_t Qpad = (Q + WQ - 1) / WQ * WQ;
    int rc = emo_ensure(ctx, (void **)&ctx->qvec, &ctx->qvec_cap, (size_t)Qpad * words * 4);
    if (rc) return rc;
    if ((rc = emo_ensure(ctx, (void **)&ctx->keys, &ctx->keys_cap, (size_t)Q * 8))) return rc;
    {
        const uint64_t total = (uint64_t)Qpad * words, blocks = (total + 255) / 256, cap = (uint64_t)ctx->sm_count * 32;
        pack_queries_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, ctx->stream>>>(src, W, bw, Q, Qpad, dim, words, WK, ctx->qvec);
        EMO_LAUNCH_CHECK(ctx);
    }
    match_init_keys_kernel<<<(Q + 255) / 256, 256, 0, ctx->stream>>>(ctx->keys, Q);
    EMO_LAUNCH_CHECK(ctx);
    const size_t smem = (size_t)64 * (WQ + WC) * 4 + 2 * 8 * 8;  // stages * WK = 64 words deep in every variant
    auto kern = WK == 32 ? match_wide_kernel<32> : (WK == 16 ? match_wide_kernel<16> : match_wide_kernel<8>);
    EMO_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t qtiles = Qpad / WQ, n_ctiles = ctx->n_chunks;
    uint32_t splits = 1;
    const uint32_t want = (uint32_t)ctx->sm_count * 6;  // 2 CTAs per SM, >= 3 waves
    if (qtiles < want) splits = (want + qtiles - 1) / qtiles;
    if (splits > n_ctiles) splits = n_ctiles;
    if (splits > 65535) splits = 65535;
    const uint32_t tps = (n_ctiles + splits - 1) / splits;
    splits = (n_ctiles + tps - 1) / tps;
    kern<<<dim3(qtiles, splits), 288, smem, ctx->stream>>>(ctx->qvec, ctx->cand, n_slices, n_ctiles, tps, Q, ctx->keys);
    EMO_LAUNCH_CHECK(ctx);
    match_finalize_kernel<<<(Q + 255) / 256, 256, 0, ctx->stream>>>(ctx->keys, Q, 1u, item, dist);
    EMO_LAUNCH_CHECK(ctx);
    return EMO_OK;
}

// 1to1: answer from the colour-cube index (index.cu) when it exists or is worth building for `queries` blocks — the
// build plus the lookup cost about as much as scanning 2^31 (block, tile) pairs.  Same results either way.
int emo_prepare_match(emo_ctx *ctx, uint64_t queries) {
    if (ctx->match_mode == EMO_MATCH_SCAN || ctx->lut_valid || !emo_index_supported(ctx)) return EMO_OK;
    if (ctx->match_mode >= EMO_MATCH_INDEX || queries * ctx->L >= (1ull << 31)) return emo_launch_build_index(ctx);
    return EMO_OK;
}

int emo_launch_match(emo_ctx *ctx, const uint8_t *src, uint32_t W, uint32_t H, int32_t *item, uint32_t *dist) {
    if (ctx->wide) return launch_match_wide(ctx, src, W, H, item, dist);
    int rc = emo_prepare_match(ctx, (uint64_t)(W / ctx->dim) * (H / ctx->dim));
    if (rc) return rc;
    if (ctx->match_mode != EMO_MATCH_SCAN && ctx->lut_valid) return emo_launch_match_index(ctx, src, W, H, item, dist);
    MatchParams p;
    p.cand = ctx->cand;
    p.chunk = ctx->chunk;
    p.n_chunks = ctx->n_chunks;
    p.chunks_per_split = ctx->n_chunks;
    p.last_len = ctx->chunk;
    p.publish_all = 0;
    p.split_ctas = 0; p.splits = 1; p.split_tile0 = 0;
    p.src = src;
    p.W = W;
    p.dim = ctx->dim;
    p.bw = W / ctx->dim;
    p.Q = p.bw * (H / ctx->dim);
    p.mirrored = ctx->N != 1;
    p.item = item;
    p.dist = dist;
    p.keys = nullptr;
    const uint32_t Q = p.Q;
    // Launch shape from tools/sweep_match.py (1to1): R = 8 queries per thread (2048-query CTAs) amortises the
    // candidate loads best once the scan is long (L >= 16k) and there are enough queries; short libraries with many
    // queries prefer R = 4 (512-query CTAs); everything else R = 2 (256-query CTAs) to keep the GPU busy.
    int shape = 2;
    if (ctx->words == 1 && Q >= 65536u) shape = ctx->L >= 16384u ? 8 : 4;
    if (const char *e = getenv("EMO_MATCH_R")) shape = atoi(e);  // tuning override
    if (ctx->words == 1 && shape == 4) return launch_match_t<1, 4, 128>(ctx, p, Q);
    const bool big = shape >= 8;
    if (ctx->words == 3) {
        // tuning override: EMO_MATCH_SHAPE3 = R * 1000 + NT
        static const int shape3 = getenv("EMO_MATCH_SHAPE3") ? atoi(getenv("EMO_MATCH_SHAPE3")) : 0;
        switch (shape3) {
            case 2128: return launch_match_t<3, 2, 128>(ctx, p, Q);
            case 4064: return launch_match_t<3, 4, 64>(ctx, p, Q);
            case 4128: return launch_match_t<3, 4, 128>(ctx, p, Q);
            case 8064: return launch_match_t<3, 8, 64>(ctx, p, Q);
            case 8128: return launch_match_t<3, 8, 128>(ctx, p, Q);
            case 8256: return launch_match_t<3, 8, 256>(ctx, p, Q);
            // variants of the <3, 2, 128> shape: window length, register cap, unroll
            case 1: if (p.chunk % 256 == 0) return launch_match_t<3, 2, 128, 256, 0, 8>(ctx, p, Q); break;
            case 2: return launch_match_t<3, 2, 128, 128, 0, 4>(ctx, p, Q);
            case 3: return launch_match_t<3, 2, 128, 128, 0, 16>(ctx, p, Q);
            case 4: return launch_match_t<3, 2, 128, 128, 7, 8>(ctx, p, Q);
            case 5: return launch_match_t<3, 2, 128, 128, 5, 8>(ctx, p, Q);
            case 6: if (p.chunk % 256 == 0) return launch_match_t<3, 2, 128, 256, 7, 8>(ctx, p, Q); break;
            case 7: if (p.chunk % 256 == 0) return launch_match_t<3, 2, 128, 256, 5, 16>(ctx, p, Q); break;
            default: break;
        }
    }
    switch (ctx->words) {
        case 1: return big ? launch_match_t<1, 8, 256>(ctx, p, Q) : launch_match_t<1, 2, 128>(ctx, p, Q);
        // 4to1: R = 2 beats every larger register tile (tools/sweep_match3.py: 0.78 vs 0.67-0.76 of the VABSDIFF4 rate on C2); 16x unroll +1.5 %
        case 3: return big ? launch_match_t<3, 8, 256>(ctx, p, Q) : launch_match_t<3, 2, 128, MATCH_WIN, 0, 16>(ctx, p, Q);
        case 7: return launch_match_t<7, 2, 128>(ctx, p, Q);
        case 12: return launch_match_t<12, 2, 128>(ctx, p, Q);
    }
    emo_set_error("match: unsupported vector length (words=%u)", ctx->words);
    return EMO_ERR_UNSUPPORTED;
}
