// topk.cu — ranked candidate lists for the no-repeat renderer.
//
// Reference: the Scoring phase of render_nto1_no_repeat, src/mosaic/rendering.rs:307-321 — for every block
// `kdtree.nearest_n::<Manhattan>(&coords, 100000)`, i.e. (practically) ALL candidates of the search set sorted by
// L1 distance, and the refill `compute_nearest(n, 10)` at :384-386.  The greedy assignment that consumes the lists
// (:341-392) is sequential and stays on the host (host/emosaic.cpp, api.py); it rarely looks past the first few
// entries of a list, so this kernel serves the lists in pages: for every block the candidates at positions
// [first, first + k) of its list ordered by (distance, insertion rank).  Insertion rank = 2t for tile t, 2t + 1 for its
// mirror (tileset.rs:178-190); kiddo's own order among equal distances is unpinned and this is the
// canonical one the oracle uses.  `exclude` ([T] bytes, optional) removes the candidates of retired tiles from every
// list — the reference removes a placed tile from the tree (:366-380), so its refill sees the pruned set.  For N == 1
// a tile and its mirror have the same vector: the mirror sits directly behind the tile in every list and using one
// retires both (rendering.rs:357-358), so it is omitted.
//
// One CTA per block.  The block's query vector stays in registers; candidates are read from the packed array built
// by emo_set_library (L2-resident).  Keys are (distance << 31 | rank), unique per candidate, so the k-th key is found
// exactly by a 4-digit radix select (12-bit digits, shared-memory histogram per digit, distances recomputed on the
// fly: 3N-byte vectors, one VABSDIFF4 per word) for both ends of the page; a last pass collects the keys in between
// (exactly the page), a bitonic sort orders them.  9 passes over L candidates per block — the lists the reference
// builds cost a full sort of L per block.
#include "common.cuh"

static constexpr int TOPK_THREADS = 256;
static constexpr int TOPK_MAX_K = 1024;
static constexpr int TOPK_BINS = 4096;  // 12-bit digits

struct TopkParams {
    const uint32_t *cand;  // [Lpad][WORDS]
    uint32_t L;            // real candidates
    uint32_t mirrored;     // 1: rank c -> tile c >> 1, mirrored c & 1; 0: rank c -> tile c
    const uint8_t *src;    // [H][W][3]
    uint32_t W, dim, bw, N;
    uint32_t first, k;
    const uint8_t *exclude;  // [T] or nullptr: 1 = tile retired (both orientations), its candidates do not exist
    int32_t *item;   // [Q][k]
    uint32_t *dist;  // [Q][k]
};

template <bool EXCL>
__device__ __forceinline__ bool topk_alive(const TopkParams &p, uint32_t c) {
    if (!EXCL) return true;
    return __ldg(p.exclude + (p.mirrored ? (c >> 1) : c)) == 0;
}

template <int WORDS>
__device__ __forceinline__ unsigned long long topk_key(const uint32_t (&q)[WORDS], const uint32_t *__restrict__ cand, uint32_t c) {
    uint32_t d = 0;
#pragma unroll
    for (int w = 0; w < WORDS; w++) d = sad4(q[w], __ldg(cand + (size_t)c * WORDS + w), d);
    return (unsigned long long)d << 31 | c;
}

// the key at sorted position `pos` (0-based) among the L keys; every thread returns it
template <int WORDS, bool EXCL>
__device__ unsigned long long topk_select(const uint32_t (&q)[WORDS], const TopkParams &p, uint32_t pos, uint32_t *hist, uint32_t *scratch) {
    unsigned long long prefix = 0, mask = 0;
    uint32_t remaining = pos;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll 1
    for (int shift = 36; shift >= 0; shift -= 12) {
        for (int b = tid; b < TOPK_BINS; b += TOPK_THREADS) hist[b] = 0;
        __syncthreads();
        for (uint32_t c = tid; c < p.L; c += TOPK_THREADS) {
            if (!topk_alive<EXCL>(p, c)) continue;
            const unsigned long long key = topk_key<WORDS>(q, p.cand, c);
            if ((key & mask) == prefix) atomicAdd(&hist[(uint32_t)(key >> shift) & (TOPK_BINS - 1)], 1u);
        }
        __syncthreads();
        // 16 consecutive bins per thread; exclusive scan of the per-thread sums over the block
        uint32_t mine = 0;
#pragma unroll
        for (int j = 0; j < TOPK_BINS / TOPK_THREADS; j++) mine += hist[tid * (TOPK_BINS / TOPK_THREADS) + j];
        uint32_t incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) scratch[warp] = incl;
        __syncthreads();
        uint32_t base = 0;
        for (int w = 0; w < warp; w++) base += scratch[w];
        const uint32_t excl = base + incl - mine;
        if (remaining >= excl && remaining < excl + mine) {  // exactly one thread: the digit lives in my 16 bins
            uint32_t acc = excl;
            for (int j = 0; j < TOPK_BINS / TOPK_THREADS; j++) {
                const uint32_t h = hist[tid * (TOPK_BINS / TOPK_THREADS) + j];
                if (remaining < acc + h) {
                    scratch[8] = tid * (TOPK_BINS / TOPK_THREADS) + j;
                    scratch[9] = remaining - acc;
                    break;
                }
                acc += h;
            }
        }
        __syncthreads();
        prefix |= (unsigned long long)scratch[8] << shift;
        mask |= (unsigned long long)(TOPK_BINS - 1) << shift;
        remaining = scratch[9];
        __syncthreads();
    }
    return prefix;
}

template <int WORDS, bool EXCL>
__global__ void __launch_bounds__(TOPK_THREADS) topk_kernel(const TopkParams p) {
    __shared__ uint32_t hist[TOPK_BINS];
    __shared__ unsigned long long keys[TOPK_MAX_K];
    __shared__ uint32_t scratch[12];
    __shared__ uint32_t qs[WORDS];
    const int tid = threadIdx.x;
    const uint32_t blk = blockIdx.x, by = blk / p.bw, bx = blk % p.bw;
    // query vector (analysis.rs:23-36): the dim x dim pixels of the block, cells row-major, r,g,b per cell
    if (tid < WORDS) {
        uint32_t packed = 0;
        for (uint32_t j = 0; j < 4; j++) {
            const uint32_t b = tid * 4 + j;
            if (b < 3 * p.N) {
                const uint32_t cell = b / 3, ch = b % 3, row = cell / p.dim, col = cell % p.dim;
                packed |= (uint32_t)p.src[((size_t)(by * p.dim + row) * p.W + bx * p.dim + col) * 3 + ch] << (8 * j);
            }
        }
        qs[tid] = packed;
    }
    __syncthreads();
    uint32_t q[WORDS];
#pragma unroll
    for (int w = 0; w < WORDS; w++) q[w] = qs[w];

    int32_t *item = p.item + (size_t)blk * p.k;
    uint32_t *dist = p.dist + (size_t)blk * p.k;
    // candidates that exist for this call (all of them, or those of the tiles not retired yet)
    uint32_t alive = p.L;
    if (EXCL) {
        if (tid == 0) scratch[11] = 0;
        __syncthreads();
        uint32_t mine = 0;
        for (uint32_t c = tid; c < p.L; c += TOPK_THREADS) mine += topk_alive<EXCL>(p, c) ? 1u : 0u;
        atomicAdd(&scratch[11], mine);
        __syncthreads();
        alive = scratch[11];
    }
    uint32_t n = 0;  // entries of this page that exist
    if (p.first < alive) n = min(p.k, alive - p.first);
    if (n > 0) {
        const unsigned long long lo = topk_select<WORDS, EXCL>(q, p, p.first, hist, scratch);
        const unsigned long long hi = n > 1 ? topk_select<WORDS, EXCL>(q, p, p.first + n - 1, hist, scratch) : lo;
        if (tid == 0) scratch[10] = 0;
        uint32_t m = 1;
        while (m < n) m <<= 1;
        for (uint32_t j = tid; j < m; j += TOPK_THREADS) keys[j] = ~0ull;
        __syncthreads();
        for (uint32_t c = tid; c < p.L; c += TOPK_THREADS) {
            if (!topk_alive<EXCL>(p, c)) continue;
            const unsigned long long key = topk_key<WORDS>(q, p.cand, c);
            if (key >= lo && key <= hi) keys[atomicAdd(&scratch[10], 1u)] = key;  // exactly n keys qualify
        }
        __syncthreads();
        for (uint32_t size = 2; size <= m; size <<= 1) {  // bitonic sort, ascending
            for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
                for (uint32_t j = tid; j < m; j += TOPK_THREADS) {
                    const uint32_t partner = j ^ stride;
                    if (partner > j) {
                        const bool up = (j & size) == 0;
                        const unsigned long long a = keys[j], b = keys[partner];
                        if ((a > b) == up) {
                            keys[j] = b;
                            keys[partner] = a;
                        }
                    }
                }
                __syncthreads();
            }
        }
    }
    for (uint32_t j = tid; j < p.k; j += TOPK_THREADS) {
        if (j < n) {
            const unsigned long long key = keys[j];
            const uint32_t c = (uint32_t)(key & 0x7FFFFFFFu);
            const uint32_t t = p.mirrored ? (c >> 1) : c;
            item[j] = (p.mirrored && (c & 1)) ? -(int32_t)(t + 1) : (int32_t)(t + 1);
            dist[j] = (uint32_t)(key >> 31);
        } else {
            item[j] = 0;
            dist[j] = 0xFFFFFFFFu;
        }
    }
}

int emo_launch_topk(emo_ctx *ctx, const uint8_t *src, uint32_t W, uint32_t H, uint32_t first, uint32_t k, const uint8_t *exclude,
                    int32_t *item, uint32_t *dist) {
    TopkParams p;
    p.cand = ctx->cand;
    p.L = ctx->L;
    p.mirrored = ctx->N != 1;
    p.src = src;
    p.W = W;
    p.dim = ctx->dim;
    p.bw = W / ctx->dim;
    p.N = ctx->N;
    p.first = first;
    p.k = k;
    p.exclude = exclude;
    p.item = item;
    p.dist = dist;
    const uint32_t Q = p.bw * (H / ctx->dim);
    const bool ex = exclude != nullptr;
    switch (ctx->words) {
        case 1: ex ? topk_kernel<1, true><<<Q, TOPK_THREADS, 0, ctx->stream>>>(p) : topk_kernel<1, false><<<Q, TOPK_THREADS, 0, ctx->stream>>>(p); break;
        case 3: ex ? topk_kernel<3, true><<<Q, TOPK_THREADS, 0, ctx->stream>>>(p) : topk_kernel<3, false><<<Q, TOPK_THREADS, 0, ctx->stream>>>(p); break;
        case 7: ex ? topk_kernel<7, true><<<Q, TOPK_THREADS, 0, ctx->stream>>>(p) : topk_kernel<7, false><<<Q, TOPK_THREADS, 0, ctx->stream>>>(p); break;
        case 12: ex ? topk_kernel<12, true><<<Q, TOPK_THREADS, 0, ctx->stream>>>(p) : topk_kernel<12, false><<<Q, TOPK_THREADS, 0, ctx->stream>>>(p); break;
        default:
            emo_set_error("topk: ranked lists exist for --mode 1..4 (N = 1, 4, 9, 16), not N=%u", ctx->N);
            return EMO_ERR_UNSUPPORTED;
    }
    EMO_LAUNCH_CHECK(ctx);
    return EMO_OK;
}
