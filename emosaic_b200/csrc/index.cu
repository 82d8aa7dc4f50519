// index.cu — the 1to1 search index: an exact L1 nearest-tile table over the 24-bit colour cube.
//
// Reference: TileSet::build_kiddo() src/mosaic/tiles/tileset.rs:178-190 builds a KD-tree over the
// library once per render (rendering.rs:136) and every block then asks it for nearest_one::<Manhattan>
// (rendering.rs:187-195).  For --mode 1 the query is one RGB pixel, so the whole query space has only
// 2^24 points: this file builds the answer for all of them and turns the per-pixel search into one
// 4-byte load.  The answers are the brute-force kernel's answers bit for bit (match.cu):
// minimum L1 distance, ties to the smallest tile index (the mirrored twin of a 1x1 colour vector is
// the vector itself and can never win, so every item is positive).
//
// Construction = exact city-block distance transform.  |dr|+|dg|+|db| separates per axis, so
//     best(r,g,b) = min_b' ( |b-b'| + min_g' ( |g-g'| + min_r' ( |r-r'| + seed(r',g',b') ) ) )
// where seed is 0 at library colours and +inf elsewhere.  Cells hold key = dist << 22 | tile, so an
// unsigned min is the lexicographic (dist, tile) min the tie-break asks for and adding k << 22 moves
// only the distance.  Each axis pass is the classic forward/backward sweep (key[i] = min(key[i],
// key[i-1] + 1<<22)), exact for L1.  dist <= 765 < 2^10 and tile < 2^22; larger libraries keep using
// the scan kernel.
//
// Layout in HBM: lut[b][g][r] u32 (cell = r | g << 8 | b << 16, i.e. the low 24 bits of the pixel
// read as a little-endian word), 64 MiB.  Cost: build = 1 memset + 4 small kernels (about 0.1 ms, once per
// library); lookup = 3 B read + 8 B written per query plus one gather.
//
// Compact form (lut16, 32 MiB): the 64 MiB table is more than the ~63 MB of L2 that one die of the B200 keeps of
// a table every SM reads, so next to a streaming output it is re-fetched from DRAM on every pass.  A cell only has
// to name its winner: lut16[cell] = slot (u16), and the distance is recomputed from the winner's colour
// (one VABSDIFF4).  slot = tile index when T <= 65 536 — every library the reference can load (i16 ids,
// tileset.rs:182) — and then the colours are the candidate array itself; larger libraries are compacted to their
// distinct winners (tiles with duplicate colours never win: C4's 100 000 tiles have 43 909 distinct colours) with an
// 8-byte {tile, colour} entry per slot.  The second lookup goes to an array of at most 512 KB (L1 / L2).
#include <stdlib.h>

#include "common.cuh"

static constexpr uint32_t IDX_TILE_BITS = EMO_IDX_TILE_BITS;
static constexpr uint32_t IDX_INC = 1u << IDX_TILE_BITS;
static constexpr uint32_t IDX_TILE_MASK = EMO_IDX_TILE_MASK;
static constexpr uint32_t IDX_EMPTY = 0xFFFFFFFFu;
static constexpr size_t IDX_CELLS = (size_t)1 << 24;

// key + k steps, keeping "no tile yet" as it is (a real key never wraps: dist <= 765, k <= 255)
__device__ __forceinline__ uint32_t idx_step(uint32_t key, uint32_t k) {
    const uint32_t n = key + k * IDX_INC;
    return n < key ? IDX_EMPTY : n;
}

// seeds: lut[colour of tile t] = min(t) — duplicates of a colour keep the smallest tile index
__global__ void index_seed_kernel(const uint32_t *__restrict__ cand, uint32_t T, uint32_t *__restrict__ lut) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < T) atomicMin(&lut[cand[t] & 0xFFFFFFu], t);
}

// The same without clearing the 64 MiB table first: a bit per cell says "a library colour lives here" (2 MiB, cleared per build),
// the marked cells are reset by this kernel, index_seed_kernel then takes the minimum over them, and the r pass reads nothing but
// the bits and the marked cells and writes every line from scratch.
__global__ void index_mark_kernel(const uint32_t *__restrict__ cand, uint32_t T, uint32_t *__restrict__ seeded, uint32_t *__restrict__ lut) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const uint32_t c = cand[t] & 0xFFFFFFu;
    atomicOr(&seeded[c >> 5], 1u << (c & 31));
    lut[c] = IDX_EMPTY;  // every writer of a cell stores the same value
}

// Pass along r (the contiguous axis): one warp per 256-cell line, 8 consecutive cells per lane (two
// LDG.128 / STG.128), prefix- and suffix-min across lanes by shuffle.  Offsetting a lane's boundary
// key by its distance to the far end of the line turns "min of key + distance" into a plain min-scan.
// one 32-byte sector per lane in a single instruction (sm_100: 256-bit global stores)
__device__ __forceinline__ void stg_v8(uint32_t *p, const uint32_t (&v)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
                 "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}

// SPARSE: the table holds valid keys only where `seeded` has a bit (index_mark_kernel); everything else counts as empty.
template <bool SPARSE>
__global__ void __launch_bounds__(256) index_sweep_r_kernel(uint32_t *__restrict__ lut, const uint32_t *__restrict__ seeded) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t line = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // g | b << 8
    uint4 *p = reinterpret_cast<uint4 *>(lut + ((size_t)line << 8)) + lane * 2;
    uint32_t k[8];
    if (SPARSE) {
        const uint32_t bits = (__ldg(seeded + line * 8 + (lane >> 2)) >> (8 * (lane & 3))) & 0xFFu;  // my 8 cells
#pragma unroll
        for (int m = 0; m < 8; m++) k[m] = IDX_EMPTY;
        if (__ballot_sync(0xFFFFFFFFu, bits != 0) == 0) {  // no library colour on this line (most lines of a clustered library)
            stg_v8(reinterpret_cast<uint32_t *>(p), k);
            return;
        }
        if (bits) {
            const uint4 a = p[0], b = p[1];
            const uint32_t v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
            for (int m = 0; m < 8; m++)
                if (bits >> m & 1) k[m] = v[m];
        }
    } else {
        const uint4 a = p[0], b = p[1];
        k[0] = a.x; k[1] = a.y; k[2] = a.z; k[3] = a.w; k[4] = b.x; k[5] = b.y; k[6] = b.z; k[7] = b.w;
    }
    uint32_t f[8], s[8];
    // inside the lane
    f[0] = k[0];
#pragma unroll
    for (int m = 1; m < 8; m++) f[m] = min(k[m], idx_step(f[m - 1], 1));
    s[7] = k[7];
#pragma unroll
    for (int m = 6; m >= 0; m--) s[m] = min(k[m], idx_step(s[m + 1], 1));
    // across lanes: forward carries leave from cell 7 of a lane, backward carries from cell 0
    uint32_t uf = idx_step(f[7], (31 - lane) * 8);  // distance to the line's last lane, in lanes * 8
    uint32_t ub = idx_step(s[0], lane * 8);
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t of = __shfl_up_sync(0xFFFFFFFFu, uf, d), ob = __shfl_down_sync(0xFFFFFFFFu, ub, d);
        if (lane >= (uint32_t)d) uf = min(uf, of);
        if (lane + d < 32) ub = min(ub, ob);
    }
    uint32_t ef = __shfl_up_sync(0xFFFFFFFFu, uf, 1), eb = __shfl_down_sync(0xFFFFFFFFu, ub, 1);
    if (lane == 0) ef = IDX_EMPTY;
    if (lane == 31) eb = IDX_EMPTY;
    // carry at my cell 0 (forward) / cell 7 (backward): remove the offset, the lane boundary is 1 cell away
    // ef = key + (8 * (lane - l') + (31 - lane) * 8) steps seen from cell 7 of lane l'; cell 0 of mine is 8 * (lane - l') - 7 away
    const uint32_t cf = ef == IDX_EMPTY ? IDX_EMPTY : ef - ((31 - lane) * 8 + 7) * IDX_INC;
    const uint32_t cb = eb == IDX_EMPTY ? IDX_EMPTY : eb - (lane * 8 + 7) * IDX_INC;
    uint32_t o[8];
#pragma unroll
    for (int m = 0; m < 8; m++) o[m] = min(min(f[m], s[m]), min(idx_step(cf, m), idx_step(cb, 7 - m)));
    stg_v8(reinterpret_cast<uint32_t *>(p), o);
}

// Pass along g (stride 256 cells) or b (stride 65 536 cells): one thread per line, neighbouring threads on
// neighbouring r so every access of a warp is one 128-byte row.  Forward sweep in place, then the backward
// sweep over the forward result (min(F, B) in one go).  Loads do not depend on the running minimum, so they
// are issued 8 deep.
template <int AXIS>  // 1: g, 2: b
__global__ void __launch_bounds__(256) index_sweep_kernel(uint32_t *__restrict__ lut) {
    const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;  // 65 536 lines
    const uint32_t r = id & 255, o = id >> 8;
    const size_t base = AXIS == 1 ? ((size_t)o << 16 | r) : ((size_t)o << 8 | r);
    const size_t stride = AXIS == 1 ? 256 : 65536;
    uint32_t *p = lut + base;
    uint32_t run = IDX_EMPTY;
    for (int i = 0; i < 256; i += 8) {
        uint32_t k[8];
#pragma unroll
        for (int m = 0; m < 8; m++) k[m] = p[(size_t)(i + m) * stride];
#pragma unroll
        for (int m = 0; m < 8; m++) {
            run = min(k[m], idx_step(run, 1));
            p[(size_t)(i + m) * stride] = run;
        }
    }
    run = IDX_EMPTY;
    for (int i = 255; i >= 0; i -= 8) {
        uint32_t k[8];
#pragma unroll
        for (int m = 0; m < 8; m++) k[m] = p[(size_t)(i - m) * stride];
#pragma unroll
        for (int m = 0; m < 8; m++) {
            run = min(k[m], idx_step(run, 1));
            p[(size_t)(i - m) * stride] = run;
        }
    }
}

// The same pass with the line in registers: one read and one write of the table per axis instead of two of each.  A CTA of
// 8 warps takes a slab of 256 positions along the axis x 32 neighbouring r (every access of a warp is one 128-byte row); warp w
// owns positions 32w .. 32w + 31 of all 32 lines, a thread its 32 keys of one line.  Inside a thread the forward and backward
// sweeps run over registers; between warps only the two boundary keys travel (shared memory, 2 x 8 x 32 words): the forward
// carry into a segment is the minimum over the earlier segments of their local end value + the distance, which is exactly what
// a sweep over the whole line would have brought along.
template <int AXIS>  // 1: g, 2: b
__global__ void __launch_bounds__(256) index_sweep_reg_kernel(uint32_t *__restrict__ lut) {
    __shared__ uint32_t carry[2][8][32];
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t r = (blockIdx.x & 7) * 32 + lane, o = blockIdx.x >> 3;
    const size_t base = AXIS == 1 ? ((size_t)o << 16 | r) : ((size_t)o << 8 | r);
    const size_t stride = AXIS == 1 ? 256 : 65536;
    uint32_t *p = lut + base + (size_t)(32 * w) * stride;
    uint32_t k[32], m[32];
#pragma unroll
    for (int j = 0; j < 32; j++) k[j] = p[(size_t)j * stride];
    m[0] = k[0];
#pragma unroll
    for (int j = 1; j < 32; j++) m[j] = min(k[j], idx_step(m[j - 1], 1));
    carry[0][w][lane] = m[31];  // local forward value at the segment's last position
    uint32_t run = k[31];
#pragma unroll
    for (int j = 30; j >= 0; j--) {
        run = min(k[j], idx_step(run, 1));
        m[j] = min(m[j], run);
    }
    carry[1][w][lane] = run;    // local backward value at the segment's first position
    __syncthreads();
    uint32_t cf = IDX_EMPTY, cb = IDX_EMPTY;  // carries as seen at my position 0 (forward) / 31 (backward)
#pragma unroll
    for (int v = 0; v < 8; v++) {
        if ((uint32_t)v < w) cf = min(cf, idx_step(carry[0][v][lane], 32 * (w - v) - 31));
        if ((uint32_t)v > w) cb = min(cb, idx_step(carry[1][v][lane], 32 * (v - w) - 31));
    }
#pragma unroll
    for (int j = 0; j < 32; j++) p[(size_t)j * stride] = min(m[j], min(idx_step(cf, j), idx_step(cb, 31 - j)));
}

// The b pass with the compact form written on the way out (MODE 1: slot = tile index, MODE 2: slot_of_tile).  Same sweep as
// index_sweep_reg_kernel<2>, but a warp spans 8 r x 4 g (four 32-byte sectors per access) so that four consecutive b of a thread
// complete whole 4 x 4 x 4 blocks of the compact layout inside the warp: the 16 lanes of an r-block write one 32-byte run of the
// block's 128-byte line per b.  Saves the compaction pass (a second read of the 64 MiB table and a launch).
__host__ __device__ __forceinline__ uint32_t idx16_pos(uint32_t c);
template <int MODE>
__global__ void __launch_bounds__(256) index_sweep_b_compact_kernel(uint32_t *__restrict__ lut, const uint32_t *__restrict__ slot_of_tile,
                                                                    uint16_t *__restrict__ lut16) {
    __shared__ uint32_t carry[2][8][32];
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t r = (blockIdx.x & 31) * 8 + (lane & 7), g = (blockIdx.x >> 5) * 4 + (lane >> 3);
    const uint32_t cell0 = (32 * w) << 16 | g << 8 | r;  // position 0 of my segment
    uint32_t *p = lut + cell0;
    const size_t stride = 65536;
    uint32_t k[32], m[32];
#pragma unroll
    for (int j = 0; j < 32; j++) k[j] = p[(size_t)j * stride];
    m[0] = k[0];
#pragma unroll
    for (int j = 1; j < 32; j++) m[j] = min(k[j], idx_step(m[j - 1], 1));
    carry[0][w][lane] = m[31];
    uint32_t run = k[31];
#pragma unroll
    for (int j = 30; j >= 0; j--) {
        run = min(k[j], idx_step(run, 1));
        m[j] = min(m[j], run);
    }
    carry[1][w][lane] = run;
    __syncthreads();
    uint32_t cf = IDX_EMPTY, cb = IDX_EMPTY;
#pragma unroll
    for (int v = 0; v < 8; v++) {
        if ((uint32_t)v < w) cf = min(cf, idx_step(carry[0][v][lane], 32 * (w - v) - 31));
        if ((uint32_t)v > w) cb = min(cb, idx_step(carry[1][v][lane], 32 * (v - w) - 31));
    }
    // position of (r, g, b = 32w) in the compact table; b + 1 is 16 entries further inside a block, b + 4 the next block along b
    const uint32_t pos0 = idx16_pos(cell0);
#pragma unroll
    for (int j = 0; j < 32; j++) {
        const uint32_t key = min(m[j], min(idx_step(cf, j), idx_step(cb, 31 - j)));
        p[(size_t)j * stride] = key;
        uint32_t slot = key & IDX_TILE_MASK;
        if (MODE == 2) slot = __ldg(slot_of_tile + slot);
        lut16[pos0 + ((uint32_t)(j >> 2) << 18) + ((uint32_t)(j & 3) << 4)] = (uint16_t)slot;
    }
}

// ---------------------------------------------------------------------------------------
// compact form: u16 slot per cell
// ---------------------------------------------------------------------------------------
static constexpr uint32_t IDX16_SLOTS = 65536;

// winners = tiles that own their own colour cell (dist 0, smallest index among duplicates); slots in arrival order
// (which slot a winner gets does not matter: results depend only on the slot -> {tile, colour} pairing)
__global__ void index_winners_kernel(const uint32_t *__restrict__ cand, uint32_t T, const uint32_t *__restrict__ lut,
                                     uint32_t *__restrict__ slot_of_tile, uint2 *__restrict__ entry, uint32_t *__restrict__ counter) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const uint32_t c = cand[t] & 0xFFFFFFu;
    if ((lut[c] & IDX_TILE_MASK) != t) return;
    const uint32_t s = atomicAdd(counter, 1u);
    if (s < IDX16_SLOTS) {
        slot_of_tile[t] = s;
        entry[s] = make_uint2(t, c);
    }
}

// Position of a colour in the compact table: blocked, 4 x 4 x 4 colours per 128-byte line (64 u16) — the pixels of a photograph
// differ from their neighbours by a few levels in EVERY channel, so a line that spans 64 levels of r at fixed (g, b) is shared by
// far fewer lanes of a warp than a line that spans 4 levels of each; a uniformly random source does not care.
__host__ __device__ __forceinline__ uint32_t idx16_pos(uint32_t c) {  // c = r | g << 8 | b << 16
    const uint32_t blk = ((c >> 2) & 63u) | ((c >> 10) & 63u) << 6 | ((c >> 18) & 63u) << 12;
    return blk << 6 | (c & 3u) | ((c >> 8) & 3u) << 2 | ((c >> 16) & 3u) << 4;
}

// lut16[idx16_pos(cell)] = slot of the cell's winner.  One thread per 4 x 4 x 4 block of the cube: sixteen 16-byte runs of the
// 64 MiB table in (neighbouring threads read neighbouring runs), the block's whole 128-byte line out.  DIRECT: slot = tile index.
template <bool DIRECT>
__global__ void __launch_bounds__(256) index_compact_kernel(const uint32_t *__restrict__ lut, const uint32_t *__restrict__ slot_of_tile,
                                                            uint16_t *__restrict__ lut16) {
    const uint32_t blk = blockIdx.x * blockDim.x + threadIdx.x;  // rb | gb << 6 | bb << 12
    const uint32_t rb = blk & 63, gb = (blk >> 6) & 63, bb = blk >> 12;
    const uint4 *in = reinterpret_cast<const uint4 *>(lut + ((size_t)(bb * 4) << 16 | (size_t)(gb * 4) << 8 | (size_t)(rb * 4)));
    uint4 *out = reinterpret_cast<uint4 *>(lut16 + ((size_t)blk << 6));
#pragma unroll
    for (int b = 0; b < 4; b++) {
        uint32_t k[16];
#pragma unroll
        for (int g = 0; g < 4; g++) {
            const uint4 v = __ldg(in + ((size_t)b << 14) + ((size_t)g << 6));  // 65 536 and 256 cells further, in units of 4 cells
            k[4 * g + 0] = v.x; k[4 * g + 1] = v.y; k[4 * g + 2] = v.z; k[4 * g + 3] = v.w;
        }
#pragma unroll
        for (int m = 0; m < 16; m++) {
            k[m] &= IDX_TILE_MASK;
            if (!DIRECT) k[m] = __ldg(slot_of_tile + k[m]);
        }
        // in-block offset = r | g << 2 | b << 4: the 16 cells of one b are 32 consecutive bytes
        out[2 * b] = make_uint4(k[0] | k[1] << 16, k[2] | k[3] << 16, k[4] | k[5] << 16, k[6] | k[7] << 16);
        out[2 * b + 1] = make_uint4(k[8] | k[9] << 16, k[10] | k[11] << 16, k[12] | k[13] << 16, k[14] | k[15] << 16);
    }
}

// the tables of the index, allocated once per ctx (emo_reserve calls this ahead of the first match)
int emo_index_reserve(emo_ctx *ctx) {
    if (!ctx->lut) EMO_CK(cudaMalloc(&ctx->lut, IDX_CELLS * sizeof(uint32_t)));
    if (!ctx->lut16) EMO_CK(cudaMalloc(&ctx->lut16, IDX_CELLS * sizeof(uint16_t)));
    if (!ctx->idx_seeded) EMO_CK(cudaMalloc(&ctx->idx_seeded, IDX_CELLS / 8));
    if (ctx->T > IDX16_SLOTS) {
        int rc = emo_ensure(ctx, (void **)&ctx->idx_slot_of_tile, &ctx->idx_slot_cap, (size_t)ctx->T * 4);
        if (rc) return rc;
        if (!ctx->idx_entry) EMO_CK(cudaMalloc(&ctx->idx_entry, IDX16_SLOTS * sizeof(uint2) + 16));
    }
    return EMO_OK;
}

// winner bookkeeping of libraries above 65 536 tiles: slots for the distinct winners, the count on its way to the host
static int index_launch_winners(emo_ctx *ctx) {
    int rc;
    if ((rc = emo_ensure(ctx, (void **)&ctx->idx_slot_of_tile, &ctx->idx_slot_cap, (size_t)ctx->T * 4))) return rc;
    if (!ctx->idx_entry) EMO_CK(cudaMalloc(&ctx->idx_entry, IDX16_SLOTS * sizeof(uint2) + 16));
    uint32_t *counter = reinterpret_cast<uint32_t *>(ctx->idx_entry + IDX16_SLOTS);
    EMO_CK(cudaMemsetAsync(counter, 0, 4, ctx->stream));
    index_winners_kernel<<<(ctx->T + 255) / 256, 256, 0, ctx->stream>>>(ctx->cand, ctx->T, ctx->lut, ctx->idx_slot_of_tile, ctx->idx_entry,
                                                                        counter);
    EMO_LAUNCH_CHECK(ctx);
    // The winner count decides whether the compact table may be used (at most 65 536 slots).  It comes back asynchronously
    // (pinned word + event) and is looked at when the first lookup is launched; the compaction itself is queued right away —
    // if there turn out to be too many winners its output is simply never read — so the build has no host round trip inside.
    if (!ctx->idx_count_host) {
        EMO_CK(cudaHostAlloc((void **)&ctx->idx_count_host, 16, cudaHostAllocDefault));
        EMO_CK(cudaEventCreateWithFlags(&ctx->idx_count_ev, cudaEventDisableTiming));
    }
    EMO_CK(cudaMemcpyAsync((void *)ctx->idx_count_host, counter, 4, cudaMemcpyDeviceToHost, ctx->stream));
    EMO_CK(cudaEventRecord(ctx->idx_count_ev, ctx->stream));
    return EMO_OK;
}

int emo_launch_build_index(emo_ctx *ctx) {
    if (!ctx->lut) EMO_CK(cudaMalloc(&ctx->lut, IDX_CELLS * sizeof(uint32_t)));
    static const bool wide_only = getenv("EMO_INDEX_WIDE") && atoi(getenv("EMO_INDEX_WIDE")) != 0;  // tests / tuning: u32 table only
    static const bool sweep_old = getenv("EMO_INDEX_SWEEP") && atoi(getenv("EMO_INDEX_SWEEP")) == 0;  // A-B switch: the first build
    const bool direct = ctx->T <= IDX16_SLOTS;  // slot = tile index, colours = the candidate array
    int rc;
    ctx->lut16_mode = 0;
    if (!wide_only && !ctx->lut16) EMO_CK(cudaMalloc(&ctx->lut16, IDX_CELLS * sizeof(uint16_t)));
    if (sweep_old) {
        // memset + seeds + in-place sweeps (two reads and two writes of the table per strided axis) + a separate compaction pass
        EMO_CK(cudaMemsetAsync(ctx->lut, 0xFF, IDX_CELLS * sizeof(uint32_t), ctx->stream));
        index_seed_kernel<<<(ctx->T + 255) / 256, 256, 0, ctx->stream>>>(ctx->cand, ctx->T, ctx->lut);
        EMO_LAUNCH_CHECK(ctx);
        index_sweep_r_kernel<false><<<65536 / 8, 256, 0, ctx->stream>>>(ctx->lut, nullptr);
        EMO_LAUNCH_CHECK(ctx);
        index_sweep_kernel<1><<<256, 256, 0, ctx->stream>>>(ctx->lut);
        EMO_LAUNCH_CHECK(ctx);
        index_sweep_kernel<2><<<256, 256, 0, ctx->stream>>>(ctx->lut);
        EMO_LAUNCH_CHECK(ctx);
        ctx->lut_valid = true;
        if (wide_only) return EMO_OK;
        if (direct) {
            index_compact_kernel<true><<<(uint32_t)(IDX_CELLS / 64 / 256), 256, 0, ctx->stream>>>(ctx->lut, nullptr, ctx->lut16);
            EMO_LAUNCH_CHECK(ctx);
        } else {
            if ((rc = index_launch_winners(ctx))) return rc;
            index_compact_kernel<false><<<(uint32_t)(IDX_CELLS / 64 / 256), 256, 0, ctx->stream>>>(ctx->lut, ctx->idx_slot_of_tile, ctx->lut16);
            EMO_LAUNCH_CHECK(ctx);
        }
    } else {
        // Seeds without clearing the table (a bit per cell), the r pass writes every line from scratch, the strided passes keep
        // their lines in registers, the last one writes the compact form on its way out.  (keeping the forward sweep of a strided
        // axis in shared memory was measured in round 2: 32 KB per warp leaves 6 warps per SM, ~55 us per axis)
        // (marking costs two passes over the tiles: above ~0.5 M tiles clearing the whole table is cheaper again)
        const bool sparse = ctx->T <= 500000;
        if (sparse) {
            if (!ctx->idx_seeded) EMO_CK(cudaMalloc(&ctx->idx_seeded, IDX_CELLS / 8));
            EMO_CK(cudaMemsetAsync(ctx->idx_seeded, 0, IDX_CELLS / 8, ctx->stream));
            index_mark_kernel<<<(ctx->T + 255) / 256, 256, 0, ctx->stream>>>(ctx->cand, ctx->T, ctx->idx_seeded, ctx->lut);
            EMO_LAUNCH_CHECK(ctx);
        } else {
            EMO_CK(cudaMemsetAsync(ctx->lut, 0xFF, IDX_CELLS * sizeof(uint32_t), ctx->stream));
        }
        index_seed_kernel<<<(ctx->T + 255) / 256, 256, 0, ctx->stream>>>(ctx->cand, ctx->T, ctx->lut);
        EMO_LAUNCH_CHECK(ctx);
        // a tile wins somewhere iff it owns its own colour cell, and the sweeps never change a seeded cell: the winners are known now
        if (!wide_only && !direct && (rc = index_launch_winners(ctx))) return rc;
        if (sparse) index_sweep_r_kernel<true><<<65536 / 8, 256, 0, ctx->stream>>>(ctx->lut, ctx->idx_seeded);
        else index_sweep_r_kernel<false><<<65536 / 8, 256, 0, ctx->stream>>>(ctx->lut, nullptr);
        EMO_LAUNCH_CHECK(ctx);
        index_sweep_reg_kernel<1><<<2048, 256, 0, ctx->stream>>>(ctx->lut);
        EMO_LAUNCH_CHECK(ctx);
        if (wide_only) index_sweep_reg_kernel<2><<<2048, 256, 0, ctx->stream>>>(ctx->lut);
        else if (direct) index_sweep_b_compact_kernel<1><<<2048, 256, 0, ctx->stream>>>(ctx->lut, nullptr, ctx->lut16);
        else index_sweep_b_compact_kernel<2><<<2048, 256, 0, ctx->stream>>>(ctx->lut, ctx->idx_slot_of_tile, ctx->lut16);
        EMO_LAUNCH_CHECK(ctx);
        ctx->lut_valid = true;
        if (wide_only) return EMO_OK;
    }
    if (direct) {
        ctx->lut16_slots = ctx->T;
        ctx->lut16_mode = 1;
    } else {
        ctx->lut16_mode = 3;  // slot -> {tile, colour} entries, pending the winner count (resolved to 2 or 0 by index16_resolve)
    }
    return EMO_OK;
}

// mode 3 -> 2 (compact table usable) or 0 (more than 65 536 distinct colours: the 64 MiB table serves every lookup)
static int index16_resolve(emo_ctx *ctx) {
    if (ctx->lut16_mode != 3) return EMO_OK;
    EMO_CK(cudaEventSynchronize(ctx->idx_count_ev));
    const uint32_t winners = *ctx->idx_count_host;
    ctx->lut16_slots = winners;
    ctx->lut16_mode = winners <= IDX16_SLOTS ? 2 : 0;
    return EMO_OK;
}

bool emo_index_supported(const emo_ctx *ctx) { return ctx->N == 1 && ctx->T <= IDX_INC; }

// ---------------------------------------------------------------------------------------
// lookup: 4 consecutive pixels (12 bytes) per thread
// ---------------------------------------------------------------------------------------
template <bool SRC_WORDS, bool OUT_VEC>
__global__ void __launch_bounds__(256) match_index_kernel(const uint8_t *__restrict__ src, const uint32_t *__restrict__ lut,
                                                          uint32_t Q, int32_t *__restrict__ item, uint32_t *__restrict__ dist) {
    const uint32_t groups = (Q + 3) >> 2;
    // L2 priorities: the table is what should stay, the source and the dist map pass through once
    const uint64_t keep = l2_policy_evict_last(), drop = l2_policy_evict_first();
    grid_dependency_wait();  // the previous kernel of the stream may still be reading the item map this one rewrites
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += gridDim.x * blockDim.x) {
        const uint32_t q0 = g << 2;
        uint32_t c[4];
        if (q0 + 4 <= Q) {
            if (SRC_WORDS) {
                const uint32_t *w = reinterpret_cast<const uint32_t *>(src) + (size_t)g * 3;
                const uint32_t w0 = ldg_nc_hint_u32(w, drop), w1 = ldg_nc_hint_u32(w + 1, drop), w2 = ldg_nc_hint_u32(w + 2, drop);
                c[0] = w0 & 0xFFFFFFu;
                c[1] = __funnelshift_r(w0, w1, 24) & 0xFFFFFFu;
                c[2] = __funnelshift_r(w1, w2, 16) & 0xFFFFFFu;
                c[3] = w2 >> 8;
            } else {
                const uint8_t *s = src + (size_t)q0 * 3;
#pragma unroll
                for (int m = 0; m < 4; m++) c[m] = s[3 * m] | (uint32_t)s[3 * m + 1] << 8 | (uint32_t)s[3 * m + 2] << 16;
            }
            uint32_t k[4];
#pragma unroll
            for (int m = 0; m < 4; m++) k[m] = ldg_nc_hint_u32(lut + c[m], keep);
            if (OUT_VEC) {
                // the item map is compose's input (kept in L2); nothing on the device reads dist again
                *reinterpret_cast<uint4 *>(item + q0) = make_uint4((k[0] & IDX_TILE_MASK) + 1, (k[1] & IDX_TILE_MASK) + 1,
                                                                   (k[2] & IDX_TILE_MASK) + 1, (k[3] & IDX_TILE_MASK) + 1);
                const uint4 dv = make_uint4(k[0] >> IDX_TILE_BITS, k[1] >> IDX_TILE_BITS, k[2] >> IDX_TILE_BITS, k[3] >> IDX_TILE_BITS);
                stg_hint_v4(dist + q0, dv, drop);
            } else {
#pragma unroll
                for (int m = 0; m < 4; m++) {
                    item[q0 + m] = (int32_t)(k[m] & IDX_TILE_MASK) + 1;
                    dist[q0 + m] = k[m] >> IDX_TILE_BITS;
                }
            }
        } else {
            for (uint32_t q = q0; q < Q; q++) {
                const uint8_t *s = src + (size_t)q * 3;
                const uint32_t k = __ldg(lut + (s[0] | (uint32_t)s[1] << 8 | (uint32_t)s[2] << 16));
                item[q] = (int32_t)(k & IDX_TILE_MASK) + 1;
                dist[q] = k >> IDX_TILE_BITS;
            }
        }
    }
}

// Lookup through the compact table: 4 pixels per thread, one u16 gather each, then the winner's colour
//   MODE 1: slot = tile, colours gathered from the candidate array
//   MODE 2: slot -> {tile, colour} entries (one 8-byte load)
// and dist = |dr| + |dg| + |db| in one VABSDIFF4 (the fourth byte of both words is zero).  The second gather goes to an
// array of at most 512 KB that lives in L1 / L2: measured free next to the table gather (a shared-memory copy of the
// colours was tried and bought nothing: 107 vs 105 us at 30 000 tiles with one 1024-thread CTA per SM, 93 vs 93 us at 4096).
template <int MODE>
__global__ void __launch_bounds__(256)
    match_index16_kernel(const uint8_t *__restrict__ src, const uint16_t *__restrict__ lut16, const uint32_t *__restrict__ colors,
                         const uint2 *__restrict__ entry, uint32_t Q, int32_t *__restrict__ item, uint32_t *__restrict__ dist) {
    const uint32_t groups = Q >> 2;  // the launcher sends multiples of 4 pixels here and the tail to the wide kernel's scalar path
    const uint64_t keep = l2_policy_evict_last(), drop = l2_policy_evict_first();
    // (an explicit griddepcontrol.launch_dependents at this point was measured: 79.8 vs 79.0 us per 512-row step — nothing to gain)
    grid_dependency_wait();  // everything below may depend on the previous kernel of the stream (index build, item map readers)
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += gridDim.x * blockDim.x) {
        const uint32_t q0 = g << 2;
        const uint32_t *w = reinterpret_cast<const uint32_t *>(src) + (size_t)g * 3;
        const uint32_t w0 = ldg_nc_hint_u32(w, drop), w1 = ldg_nc_hint_u32(w + 1, drop), w2 = ldg_nc_hint_u32(w + 2, drop);
        uint32_t c[4], sl[4], it[4], d[4];
        c[0] = w0 & 0xFFFFFFu;
        c[1] = __funnelshift_r(w0, w1, 24) & 0xFFFFFFu;
        c[2] = __funnelshift_r(w1, w2, 16) & 0xFFFFFFu;
        c[3] = w2 >> 8;
#pragma unroll
        for (int m = 0; m < 4; m++) sl[m] = ldg_nc_hint_u16(lut16 + idx16_pos(c[m]), keep);
#pragma unroll
        for (int m = 0; m < 4; m++) {
            uint32_t col;
            if (MODE == 1) {
                col = __ldg(colors + sl[m]) & 0xFFFFFFu;
                it[m] = sl[m] + 1;
            } else {
                const uint2 e = __ldg(entry + sl[m]);
                col = e.y;
                it[m] = e.x + 1;
            }
            d[m] = sad4(c[m], col, 0);
        }
        *reinterpret_cast<uint4 *>(item + q0) = make_uint4(it[0], it[1], it[2], it[3]);  // compose's input: kept in L2
        stg_hint_v4(dist + q0, make_uint4(d[0], d[1], d[2], d[3]), drop);                // nothing on the device reads dist again
    }
}

template <int MODE>
static int launch_index16(emo_ctx *ctx, const uint8_t *src, uint32_t Q4, int32_t *item, uint32_t *dist) {
    const uint32_t groups = Q4 >> 2;
    // (two groups per thread, all loads in flight together, were measured: 20.0 vs 20.7 us on a 512-row stripe, 98.6 vs 98.0 us
    // on the whole image — the gather is not limited by per-thread parallelism; one group kept)
    const uint32_t blocks = min((groups + 255) / 256, (uint32_t)ctx->sm_count * 8 * 4);
    EMO_CK(emo_launch_pdl(match_index16_kernel<MODE>, dim3(blocks), dim3(256), 0, ctx->stream, src, (const uint16_t *)ctx->lut16,
                          (const uint32_t *)ctx->cand, (const uint2 *)ctx->idx_entry, Q4, item, dist));
    EMO_LAUNCH_CHECK(ctx);
    return EMO_OK;
}

static int launch_index32(emo_ctx *ctx, const uint8_t *src, uint32_t Q, int32_t *item, uint32_t *dist) {
    const uint32_t groups = (Q + 3) >> 2;
    const uint32_t cap = (uint32_t)ctx->sm_count * 8 * 4;
    const uint32_t blocks = min((groups + 255) / 256, cap);
    const bool sw = (uintptr_t)src % 4 == 0;
    const bool ov = (uintptr_t)item % 16 == 0 && (uintptr_t)dist % 16 == 0;
    static const bool pdl = !(getenv("EMO_PDL") && atoi(getenv("EMO_PDL")) == 0);  // EMO_PDL=0: plain launches
    if (sw && ov && pdl) EMO_CK(emo_launch_pdl(match_index_kernel<true, true>, dim3(blocks), dim3(256), 0, ctx->stream, src, (const uint32_t *)ctx->lut, Q, item, dist));
    else if (sw && ov) match_index_kernel<true, true><<<blocks, 256, 0, ctx->stream>>>(src, ctx->lut, Q, item, dist);
    else if (sw) match_index_kernel<true, false><<<blocks, 256, 0, ctx->stream>>>(src, ctx->lut, Q, item, dist);
    else if (ov) match_index_kernel<false, true><<<blocks, 256, 0, ctx->stream>>>(src, ctx->lut, Q, item, dist);
    else match_index_kernel<false, false><<<blocks, 256, 0, ctx->stream>>>(src, ctx->lut, Q, item, dist);
    EMO_LAUNCH_CHECK(ctx);
    return EMO_OK;
}

int emo_launch_match_index(emo_ctx *ctx, const uint8_t *src, uint32_t W, uint32_t H, int32_t *item, uint32_t *dist) {
    const uint32_t Q = W * H;  // dim == 1: the source is a flat run of Q pixels
    // The compact table whenever it exists and the buffers allow word accesses (measured on C4: 97 vs 105 us for 16.7 M pixels,
    // 20.8 vs 28.1 us for a 2 M-pixel stripe); EMO_MATCH_INDEX_WIDE forces the 64 MiB table (tests, tuning).
    const bool aligned = (uintptr_t)src % 4 == 0 && (uintptr_t)item % 16 == 0 && (uintptr_t)dist % 16 == 0;
    if (int rcr = index16_resolve(ctx)) return rcr;
    if (ctx->lut16_mode == 0 || !aligned || Q < 4 || ctx->match_mode == EMO_MATCH_INDEX_WIDE) return launch_index32(ctx, src, Q, item, dist);
    const uint32_t Q4 = Q & ~3u;
    int rc = ctx->lut16_mode == 1 ? launch_index16<1>(ctx, src, Q4, item, dist) : launch_index16<2>(ctx, src, Q4, item, dist);
    if (rc || Q4 == Q) return rc;
    return launch_index32(ctx, src + (size_t)Q4 * 3, Q - Q4, item + Q4, dist + Q4);  // the last 1..3 pixels
}
