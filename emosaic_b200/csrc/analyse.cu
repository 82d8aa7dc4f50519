// analyse.cu — tile analysis (SURVEY §8 rows A1+A2).
//
// Reference: analyse::<N>() src/mosaic/analysis.rs:5-20 calling average_color()
// src/mosaic/color.rs:14-42 for every tile (src/main.rs:786-794): split the ts x ts tile into
// dim x dim cells of floor(ts/dim)^2 pixels (row-major, edge remainder ignored), per channel
// u64 sum / count with truncating division, stored as u8.  Everything is integer, so the result
// must be bit-identical.
//
// Data layout in HBM: tiles [T][ts][ts][3] u8, contiguous (a tile is one contiguous run of
// ts*ts*3 bytes: 12 288 B at ts=64).  Outputs [T][dim*dim][3] u8.
//
// Roofline: pure HBM read stream, 3*ts*ts bytes in per tile, 3*N bytes out
// (12 288 + 15 B per tile for the fused 1to1+4to1 pass at ts=64).
//
// Fast kernels (ts = 8, 16, 32, 64): one warp owns a private ring of shared-memory stages, lane 0
// feeds it with TMA 1-D bulk copies (cp.async.bulk, completion on an mbarrier), so the only
// global-memory instructions in the kernel are a handful of bulk copies per warp and ~4 KB..6 KB
// per stage are in flight per warp (>= 144 KB per SM).  A stage is a run of 32 tile rows; lanes
// read it back as 48-byte groups (16 pixels, channel phase fixed) with three conflict-free
// LDS.128 and sum bytes per channel with IDP.4A against byte-select masks.
#include "common.cuh"

// ---------------------------------------------------------------------------------------
// generic kernel: one thread per (tile, cell); any ts / dim
// ---------------------------------------------------------------------------------------
__global__ void analyse_generic_kernel(const uint8_t *__restrict__ tiles, uint64_t T, uint32_t ts, uint32_t dim,
                                       uint8_t *__restrict__ out) {
    const uint32_t N = dim * dim;
    const uint64_t total = T * N;
    const uint32_t cell = ts / dim;  // floor
    const uint32_t count = cell * cell;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t t = i / N;
        const uint32_t c = (uint32_t)(i % N);
        const uint32_t top = (c / dim) * cell, left = (c % dim) * cell;
        const uint8_t *base = tiles + t * (uint64_t)ts * ts * 3;
        uint32_t r = 0, g = 0, b = 0;  // <= 255 * 4096^2 < 2^32
        for (uint32_t y = 0; y < cell; y++) {
            const uint8_t *p = base + ((size_t)(top + y) * ts + left) * 3;
            for (uint32_t x = 0; x < cell; x++) {
                r += p[3 * x];
                g += p[3 * x + 1];
                b += p[3 * x + 2];
            }
        }
        uint8_t *o = out + i * 3;
        o[0] = (uint8_t)(r / count);
        o[1] = (uint8_t)(g / count);
        o[2] = (uint8_t)(b / count);
    }
}

// ---------------------------------------------------------------------------------------
// fast kernel
// ---------------------------------------------------------------------------------------
template <int TS>
struct FastCfg {
    static constexpr int ROWS_PER_STAGE = 32;
    static constexpr int STAGE_BYTES = ROWS_PER_STAGE * TS * 3;     // 6144 (ts=64) / 3072 (ts=32)
    static constexpr int STAGES_PER_TILE = TS / ROWS_PER_STAGE;     // 2 / 1
    static constexpr int GROUPS_PER_LANE = STAGE_BYTES / 48 / 32;   // 4 / 2
    static constexpr int WARPS = (TS == 64) ? 8 : 16;
    static constexpr int STAGES = 4;                                 // ring depth per warp
    static constexpr int COLBIT = (TS == 64) ? 1 : 0;                // lane bit that selects the cell column
    static constexpr size_t SMEM = (size_t)WARPS * STAGES * STAGE_BYTES + WARPS * STAGES * 8 + 128;
};

__device__ __forceinline__ void sum48(const uint4 a, const uint4 b, const uint4 c, uint32_t &r, uint32_t &g, uint32_t &bl) {
    // 12 words = 16 pixels; word phase p = index % 3: p0 = (r g b r), p1 = (g b r g), p2 = (b r g b)
    const uint32_t w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
#pragma unroll
    for (int i = 0; i < 12; i++) {
        const int p = i % 3;
        const uint32_t mr = p == 0 ? 0x01000001u : (p == 1 ? 0x00010000u : 0x00000100u);
        const uint32_t mg = p == 0 ? 0x00000100u : (p == 1 ? 0x01000001u : 0x00010000u);
        const uint32_t mb = p == 0 ? 0x00010000u : (p == 1 ? 0x00000100u : 0x01000001u);
        r = __dp4a(w[i], mr, r);
        g = __dp4a(w[i], mg, g);
        bl = __dp4a(w[i], mb, bl);
    }
}

template <int TS, bool OUT1, bool OUT4>
__global__ void __launch_bounds__(FastCfg<TS>::WARPS * 32, 1)
analyse_fast_kernel(const uint8_t *__restrict__ tiles, uint64_t T, uint8_t *__restrict__ out1, uint8_t *__restrict__ out4) {
    using C = FastCfg<TS>;
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *ring = smem + (size_t)warp * C::STAGES * C::STAGE_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)C::WARPS * C::STAGES * C::STAGE_BYTES) + warp * C::STAGES;

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < C::STAGES; s++) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncwarp();

    const uint64_t gw = (uint64_t)blockIdx.x * C::WARPS + warp;   // global warp id
    const uint64_t nw = (uint64_t)gridDim.x * C::WARPS;
    const uint64_t my_tiles = T > gw ? (T - gw + nw - 1) / nw : 0;
    const uint64_t n_stage_loads = my_tiles * C::STAGES_PER_TILE;

    auto issue = [&](uint64_t n) {  // lane 0 only: bulk-load this warp's n-th stage
        const uint64_t tile = gw + (n / C::STAGES_PER_TILE) * nw;
        const uint32_t sub = (uint32_t)(n % C::STAGES_PER_TILE);
        const int s = (int)(n % C::STAGES);
        mbar_arrive_expect_tx(&bars[s], C::STAGE_BYTES);
        bulk_g2s(ring + (size_t)s * C::STAGE_BYTES, tiles + tile * (uint64_t)(TS * TS * 3) + (size_t)sub * C::STAGE_BYTES,
                 C::STAGE_BYTES, &bars[s]);
    };

    if (lane == 0)
        for (uint64_t n = 0; n < (uint64_t)(C::STAGES - 1) && n < n_stage_loads; n++) issue(n);

    uint32_t acc[2][3] = {{0, 0, 0}, {0, 0, 0}};  // [cell row][channel], this lane's cell column
    for (uint64_t n = 0; n < n_stage_loads; n++) {
        // the stage that held load n-1 was fully read before the __syncwarp() closing the previous iteration
        if (lane == 0 && n + C::STAGES - 1 < n_stage_loads) issue(n + C::STAGES - 1);
        const int s = (int)(n % C::STAGES);
        mbar_wait(&bars[s], (uint32_t)((n / C::STAGES) & 1));
        const uint8_t *st = ring + (size_t)s * C::STAGE_BYTES;
        const uint32_t sub = (uint32_t)(n % C::STAGES_PER_TILE);
#pragma unroll
        for (int i = 0; i < C::GROUPS_PER_LANE; i++) {
            const uint4 *p = reinterpret_cast<const uint4 *>(st + (size_t)(i * 32 + lane) * 48);
            const uint4 a = p[0], b = p[1], c = p[2];
            // ts=64: the whole stage is cell row `sub`; ts=32: group row = i*16 + lane/2 -> cell row i
            if (TS == 64) {
                if (sub == 0) sum48(a, b, c, acc[0][0], acc[0][1], acc[0][2]);
                else sum48(a, b, c, acc[1][0], acc[1][1], acc[1][2]);
            } else {
                sum48(a, b, c, acc[i][0], acc[i][1], acc[i][2]);
            }
        }
        __syncwarp();
        if (sub == C::STAGES_PER_TILE - 1) {
            // reduce over the lanes that share this lane's cell column (all lane bits except COLBIT)
#pragma unroll
            for (int cr = 0; cr < 2; cr++)
#pragma unroll
                for (int ch = 0; ch < 3; ch++) {
                    uint32_t v = acc[cr][ch];
#pragma unroll
                    for (int m = 1; m < 32; m <<= 1)
                        if (m != (1 << C::COLBIT)) v += __shfl_xor_sync(0xffffffffu, v, m);
                    acc[cr][ch] = v;
                }
            const uint64_t tile = gw + (n / C::STAGES_PER_TILE) * nw;
            const int col = (lane >> C::COLBIT) & 1;
            constexpr uint32_t CELL_PX = (TS / 2) * (TS / 2);
            if (OUT1) {
                uint32_t tot[3];
#pragma unroll
                for (int ch = 0; ch < 3; ch++) {
                    uint32_t v = acc[0][ch] + acc[1][ch];
                    v += __shfl_xor_sync(0xffffffffu, v, 1 << C::COLBIT);
                    tot[ch] = v / (uint32_t)(TS * TS);
                }
                if (lane == 0) {
                    uint8_t *o = out1 + tile * 3;
                    o[0] = (uint8_t)tot[0]; o[1] = (uint8_t)tot[1]; o[2] = (uint8_t)tot[2];
                }
            }
            if (OUT4 && (lane & ~(1 << C::COLBIT)) == 0) {
#pragma unroll
                for (int cr = 0; cr < 2; cr++) {
                    uint8_t *o = out4 + tile * 12 + (cr * 2 + col) * 3;
                    o[0] = (uint8_t)(acc[cr][0] / CELL_PX);
                    o[1] = (uint8_t)(acc[cr][1] / CELL_PX);
                    o[2] = (uint8_t)(acc[cr][2] / CELL_PX);
                }
            }
#pragma unroll
            for (int cr = 0; cr < 2; cr++)
#pragma unroll
                for (int ch = 0; ch < 3; ch++) acc[cr][ch] = 0;
        }
    }
}

template <int TS, bool OUT1, bool OUT4>
static int launch_fast(emo_ctx *ctx, const uint8_t *tiles, uint64_t T, uint8_t *out1, uint8_t *out4) {
    using C = FastCfg<TS>;
    auto kern = analyse_fast_kernel<TS, OUT1, OUT4>;
    EMO_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    uint64_t want = (T + C::WARPS - 1) / C::WARPS;
    int grid = (int)(want < (uint64_t)ctx->sm_count ? want : (uint64_t)ctx->sm_count);  // 1 CTA per SM (smem-bound)
    kern<<<grid, C::WARPS * 32, C::SMEM, ctx->stream>>>(tiles, T, out1, out4);
    EMO_LAUNCH_CHECK(ctx);
    return EMO_OK;
}

// ---------------------------------------------------------------------------------------
// small tiles (ts = 8, 16 — the reference's default tile size is 16): same per-warp TMA ring, but a stage
// (1536 B = 32 groups of 48 B, one group per lane) holds 8 / 2 whole tiles
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void sum12(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t &r, uint32_t &g, uint32_t &b) {
    // one 12-byte run = 4 pixels: (r g b r)(g b r g)(b r g b)
    r = __dp4a(w0, 0x01000001u, r); g = __dp4a(w0, 0x00000100u, g); b = __dp4a(w0, 0x00010000u, b);
    r = __dp4a(w1, 0x00010000u, r); g = __dp4a(w1, 0x01000001u, g); b = __dp4a(w1, 0x00000100u, b);
    r = __dp4a(w2, 0x00000100u, r); g = __dp4a(w2, 0x00010000u, g); b = __dp4a(w2, 0x01000001u, b);
}

template <int TS, bool OUT1, bool OUT4>
__global__ void __launch_bounds__(512, 1)
analyse_small_kernel(const uint8_t *__restrict__ tiles, uint64_t T, uint8_t *__restrict__ out1, uint8_t *__restrict__ out4) {
    constexpr int STAGE_BYTES = 1536, STAGES = 8, WARPS = 16;
    constexpr int TILE_BYTES = TS * TS * 3;            // 768 / 192
    constexpr int TPS = STAGE_BYTES / TILE_BYTES;      // tiles per stage: 2 / 8
    constexpr int LPT = 32 / TPS;                      // lanes per tile: 16 / 4
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool out4_words = OUT4 && (reinterpret_cast<uintptr_t>(out4) & 3) == 0;
    uint8_t *ring = smem + (size_t)warp * STAGES * STAGE_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)WARPS * STAGES * STAGE_BYTES) + warp * STAGES;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncwarp();
    const uint64_t n_stage_total = (T + TPS - 1) / TPS;                 // stages in the whole library
    const uint64_t gw = (uint64_t)blockIdx.x * WARPS + warp, nw = (uint64_t)gridDim.x * WARPS;
    const uint64_t my = n_stage_total > gw ? (n_stage_total - gw + nw - 1) / nw : 0;
    auto issue = [&](uint64_t n) {  // lane 0: bulk-load this warp's n-th stage (the last stage of the library may be short)
        const uint64_t st = gw + n * nw, first = st * TPS;
        const uint32_t bytes = (uint32_t)((T - first < (uint64_t)TPS ? T - first : (uint64_t)TPS) * TILE_BYTES);
        const int s = (int)(n % STAGES);
        mbar_arrive_expect_tx(&bars[s], bytes);
        bulk_g2s(ring + (size_t)s * STAGE_BYTES, tiles + first * TILE_BYTES, bytes, &bars[s]);
    };
    if (lane == 0)
        for (uint64_t n = 0; n < (uint64_t)(STAGES - 1) && n < my; n++) issue(n);
    for (uint64_t n = 0; n < my; n++) {
        if (lane == 0 && n + STAGES - 1 < my) issue(n + STAGES - 1);
        const int s = (int)(n % STAGES);
        mbar_wait(&bars[s], (uint32_t)((n / STAGES) & 1));
        const uint4 *p = reinterpret_cast<const uint4 *>(ring + (size_t)s * STAGE_BYTES + lane * 48);
        const uint64_t tile = (gw + n * nw) * TPS + lane / LPT;
        uint32_t L[3] = {0, 0, 0}, Rr[3] = {0, 0, 0};  // left / right cell column of this lane's 48-byte group
        if (tile < T) {  // a short last stage leaves stale bytes behind the valid tiles
            const uint4 a = p[0], b = p[1], c = p[2];
            if (TS == 16) {  // group = one 16-pixel row: pixels 0-7 | 8-15
                sum12(a.x, a.y, a.z, L[0], L[1], L[2]);  sum12(a.w, b.x, b.y, L[0], L[1], L[2]);
                sum12(b.z, b.w, c.x, Rr[0], Rr[1], Rr[2]); sum12(c.y, c.z, c.w, Rr[0], Rr[1], Rr[2]);
            } else {         // group = two 8-pixel rows: (row0: 0-3 | 4-7) (row1: 0-3 | 4-7)
                sum12(a.x, a.y, a.z, L[0], L[1], L[2]);  sum12(a.w, b.x, b.y, Rr[0], Rr[1], Rr[2]);
                sum12(b.z, b.w, c.x, L[0], L[1], L[2]);  sum12(c.y, c.z, c.w, Rr[0], Rr[1], Rr[2]);
            }
        }
        __syncwarp();
        // lanes of one cell row: TS=16 -> 8 consecutive lanes (rows), TS=8 -> 2 consecutive lanes (row pairs)
        constexpr int CR_LANES = LPT / 2;
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
#pragma unroll
            for (int m = 1; m < CR_LANES; m <<= 1) {
                L[ch] += __shfl_xor_sync(0xffffffffu, L[ch], m);
                Rr[ch] += __shfl_xor_sync(0xffffffffu, Rr[ch], m);
            }
        }
        constexpr uint32_t CELL_PX = (TS / 2) * (TS / 2);
        if (OUT4) {
            // the 6 bytes of this lane's cell row in two words; the first lane of a tile fetches the other cell row's
            // pair and writes the tile's 12 bytes as three words (a tile's output is 4-byte aligned when out4 is)
            const uint32_t w0 = (L[0] / CELL_PX) | (L[1] / CELL_PX) << 8 | (L[2] / CELL_PX) << 16 | (Rr[0] / CELL_PX) << 24;
            const uint32_t w1 = (Rr[1] / CELL_PX) | (Rr[2] / CELL_PX) << 8;
            const uint32_t x0 = __shfl_down_sync(0xffffffffu, w0, CR_LANES), x1 = __shfl_down_sync(0xffffffffu, w1, CR_LANES);
            if (tile < T && (lane % LPT) == 0) {
                if (out4_words) {
                    uint32_t *o = reinterpret_cast<uint32_t *>(out4 + tile * 12);
                    o[0] = w0;
                    o[1] = w1 | x0 << 16;
                    o[2] = x0 >> 16 | x1 << 16;
                } else {
                    uint8_t *o = out4 + tile * 12;
#pragma unroll
                    for (int k = 0; k < 4; k++) o[k] = (uint8_t)(w0 >> (8 * k));
                    o[4] = (uint8_t)w1; o[5] = (uint8_t)(w1 >> 8);
#pragma unroll
                    for (int k = 0; k < 4; k++) o[6 + k] = (uint8_t)(x0 >> (8 * k));
                    o[10] = (uint8_t)x1; o[11] = (uint8_t)(x1 >> 8);
                }
            }
        }
        if (OUT1) {
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
                uint32_t v = L[ch] + Rr[ch];
                v += __shfl_xor_sync(0xffffffffu, v, CR_LANES);  // the other cell row of the tile
                L[ch] = v / (uint32_t)(TS * TS);
            }
            if (tile < T && (lane % LPT) == 0) {
                uint8_t *o = out1 + tile * 3;
                o[0] = (uint8_t)L[0]; o[1] = (uint8_t)L[1]; o[2] = (uint8_t)L[2];
            }
        }
    }
}

template <int TS, bool OUT1, bool OUT4>
static int launch_small(emo_ctx *ctx, const uint8_t *tiles, uint64_t T, uint8_t *out1, uint8_t *out4) {
    auto kern = analyse_small_kernel<TS, OUT1, OUT4>;
    const size_t smem = (size_t)16 * 8 * 1536 + 16 * 8 * 8 + 128;
    EMO_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint64_t stages = (T * (TS * TS * 3) + 1535) / 1536, want = (stages + 15) / 16;
    const int grid = (int)(want < (uint64_t)ctx->sm_count ? want : (uint64_t)ctx->sm_count);
    kern<<<grid, 512, smem, ctx->stream>>>(tiles, T, out1, out4);
    EMO_LAUNCH_CHECK(ctx);
    return EMO_OK;
}

template <bool OUT1, bool OUT4>
static int launch_any_fast(emo_ctx *ctx, const uint8_t *tiles, uint64_t T, uint32_t ts, uint8_t *out1, uint8_t *out4) {
    switch (ts) {
        case 64: return launch_fast<64, OUT1, OUT4>(ctx, tiles, T, out1, out4);
        case 32: return launch_fast<32, OUT1, OUT4>(ctx, tiles, T, out1, out4);
        case 16: return launch_small<16, OUT1, OUT4>(ctx, tiles, T, out1, out4);
        default: return launch_small<8, OUT1, OUT4>(ctx, tiles, T, out1, out4);
    }
}

static bool fast_ok(const uint8_t *tiles, uint32_t ts) {
    return (ts == 8 || ts == 16 || ts == 32 || ts == 64) && ((uintptr_t)tiles % 16 == 0);
}

int emo_launch_analyse(emo_ctx *ctx, const uint8_t *tiles, uint64_t T, uint32_t ts, uint32_t dim, uint8_t *out) {
    if (fast_ok(tiles, ts) && (dim == 1 || dim == 2)) {
        return dim == 1 ? launch_any_fast<true, false>(ctx, tiles, T, ts, out, nullptr)
                        : launch_any_fast<false, true>(ctx, tiles, T, ts, nullptr, out);
    }
    uint64_t total = T * dim * dim;
    uint64_t blocks = (total + 255) / 256;
    uint64_t cap = (uint64_t)ctx->sm_count * 32;
    analyse_generic_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, ctx->stream>>>(tiles, T, ts, dim, out);
    EMO_LAUNCH_CHECK(ctx);
    return EMO_OK;
}

int emo_launch_analyse_fused(emo_ctx *ctx, const uint8_t *tiles, uint64_t T, uint32_t ts, uint8_t *out1, uint8_t *out4) {
    if (fast_ok(tiles, ts)) return launch_any_fast<true, true>(ctx, tiles, T, ts, out1, out4);
    int rc = emo_launch_analyse(ctx, tiles, T, ts, 1, out1);
    if (rc) return rc;
    return emo_launch_analyse(ctx, tiles, T, ts, 2, out4);
}
