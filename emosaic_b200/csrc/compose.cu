// compose.cu — output compositing fused with the tint blend.
//
// Reference: render() src/mosaic/rendering.rs:51-101 + TileSet::get_image()
// src/mosaic/tiles/tileset.rs:146-161:
//     out[by*ts + r][bx*ts + c] = tile[|item|-1][r][item < 0 ? ts-1-c : c]
// and the tint block src/main.rs:447-478: RGBA overlay of the source (alpha A), nearest-resized to
// the output, blended over the opaque mosaic with image 0.25.2 `Rgba::blend` (src-over in f32,
// truncating casts).  out_channels 3 = plain mosaic, 4 = mosaic + tint as RGBA.
//
// Layout in HBM: lib_px [2T][ts][ts][3] (entry 2t = tile t, 2t+1 = tile t mirrored, built once in
// emo_set_library) so compositing is a pure gather; item [bh][bw] i32; out [bh*ts][bw*ts][C].
// Roofline: HBM write stream, ts*ts*C bytes per placed tile; the library stays L2-resident.
//
// Bit-exact tint without per-byte f32 division: with bg alpha = 255 the f32 pipeline equals
// q = floor((fg*A + bg*(255-A)) / 255) except for a sparse exception set where the exact quotient
// is an integer and f32 rounding lands one below (q-1).  emo_prepare_tint() evaluates the exact
// f32 sequence for all 65 536 (bg, fg) pairs ON THE DEVICE (__fdiv_rn/__fmul_rn/__fadd_rn, never
// contracted), checks that claim, and stores per fg the <= 3 exceptional bg values.  The fast
// kernel then works on two 16-bit lanes per 32-bit register: X = bg*(255-A) + (fg*A + 1),
// q = (X + (X >> 8)) >> 8, and applies the exceptions branch-free by decrementing an exceptional
// bg byte before the multiply (bg-1 gives exactly q-1).  Alphas with more than 3 exceptional bg per
// fg (32 of 254) and odd geometries use the generic kernel, which reads the exact table.
#include <stdlib.h>

#include "common.cuh"

// ---------------------------------------------------------------------------------------
// tint tables
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint8_t blend_exact(uint32_t bg, uint32_t fg, uint32_t A, uint8_t *alpha_byte) {
    // image 0.25.2 src/color.rs `impl Blend for Rgba<u8>`, bg alpha = 255, fg alpha = A
    const float max_t = 255.0f;
    const float bg_a = __fdiv_rn(255.0f, max_t), fg_a = __fdiv_rn((float)A, max_t);
    const float alpha_final = __fsub_rn(__fadd_rn(bg_a, fg_a), __fmul_rn(bg_a, fg_a));
    const float b = __fdiv_rn((float)bg, max_t), f = __fdiv_rn((float)fg, max_t);
    const float b_a = __fmul_rn(b, bg_a), f_a = __fmul_rn(f, fg_a);
    const float o_a = __fadd_rn(f_a, __fmul_rn(b_a, __fsub_rn(1.0f, fg_a)));
    const float o = __fdiv_rn(o_a, alpha_final);
    if (alpha_byte) *alpha_byte = (uint8_t)__float2uint_rz(__fmul_rn(max_t, alpha_final));
    return (uint8_t)__float2uint_rz(__fmul_rn(max_t, o));
}

// grid 256 (fg), block 256 (bg)
__global__ void tint_tables_kernel(uint32_t A, uint8_t *__restrict__ lut, uint8_t *__restrict__ excv,
                                   uint8_t *__restrict__ excm, uint8_t *__restrict__ cadd, uint32_t *__restrict__ meta) {
    __shared__ uint32_t n_exc, n_r0;
    __shared__ uint8_t list[256];
    const uint32_t fg = blockIdx.x, bg = threadIdx.x;
    if (bg == 0) { n_exc = 0; n_r0 = 0; }
    __syncthreads();
    uint8_t ab = 255, v;
    if (A == 0) v = (uint8_t)bg;            // fg.a == 0: keep bg
    else if (A == 255) v = (uint8_t)fg;     // fg.a == max: copy fg
    else v = blend_exact(bg, fg, A, &ab);
    lut[bg * 256 + fg] = v;
    const uint32_t x = fg * A + bg * (255 - A);
    const uint32_t q = x / 255, rem = x % 255;
    if (rem == 0 && x > 0) atomicAdd(&n_r0, 1);
    if (v != q) {
        // the fast path relies on: exception => exact quotient integral, result q-1, bg > 0
        if (!(rem == 0 && v + 1 == q && bg > 0)) atomicAdd(&meta[2], 1);
        list[atomicAdd(&n_exc, 1)] = (uint8_t)bg;
    }
    __syncthreads();
    if (bg == 0) {
        atomicMax(&meta[0], n_exc);
        if (fg == 0) meta[1] = ab;
        // "uniform" fg: either no exception, or EVERY bg whose exact quotient is integral (x > 0) is an
        // exception.  Then q = (Y + (Y >> 8)) >> 8 with Y = x + 1 (no exception) or Y = x (all exceptions:
        // that form yields q-1 exactly at the integral quotients) needs no per-pixel fix-up at all.
        const bool uniform = n_exc == 0 || n_exc == n_r0;
        if (!uniform) atomicAdd(&meta[3], 1);
        cadd[fg] = (n_exc != 0 && uniform) ? 0 : 1;
        // deterministic order
        for (uint32_t i = 1; i < n_exc; i++)
            for (uint32_t j = i; j > 0 && list[j - 1] > list[j]; j--) {
                uint8_t t = list[j]; list[j] = list[j - 1]; list[j - 1] = t;
            }
        for (uint32_t k = 0; k < 4; k++) {
            excv[k * 256 + fg] = k < n_exc ? list[k] : 0;
            excm[k * 256 + fg] = k < n_exc ? 0x80 : 0;
        }
    }
}

int emo_prepare_tint(emo_ctx *ctx, uint8_t alpha) {
    emo_tint_tables &t = ctx->tint;
    if (t.alpha == (int)alpha) return EMO_OK;
    if (!t.lut) {
        EMO_CK(cudaMalloc(&t.lut, 65536));
        EMO_CK(cudaMalloc(&t.excv, 1024));
        EMO_CK(cudaMalloc(&t.excm, 1024));
        EMO_CK(cudaMalloc(&t.cadd, 256));
        EMO_CK(cudaMalloc(&t.meta, 16));
    }
    EMO_CK(cudaMemsetAsync(t.meta, 0, 16, ctx->stream));
    tint_tables_kernel<<<256, 256, 0, ctx->stream>>>(alpha, t.lut, t.excv, t.excm, t.cadd, t.meta);
    EMO_LAUNCH_CHECK(ctx);
    uint32_t meta[4];
    EMO_CK(cudaMemcpyAsync(meta, t.meta, 16, cudaMemcpyDeviceToHost, ctx->stream));
    EMO_CK(cudaStreamSynchronize(ctx->stream));
    t.K = (meta[2] != 0 || alpha == 0 || alpha == 255) ? 99 : (int)meta[0];  // 99: fast path not applicable
    if (t.K != 99 && meta[3] == 0) {
        t.K = 0;  // every fg is uniform: the per-fg additive constant carries the exceptions
    } else {
        EMO_CK(cudaMemsetAsync(t.cadd, 1, 256, ctx->stream));
    }
    t.alpha_out = (uint8_t)meta[1];
    t.alpha = alpha;
    return EMO_OK;
}

// ---------------------------------------------------------------------------------------
// copy kernels (RGB out)
// ---------------------------------------------------------------------------------------
template <typename V> struct VecIO;
template <> struct VecIO<uint4> {
    static __device__ __forceinline__ uint4 ld(const uint8_t *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }
    static __device__ __forceinline__ void st(uint8_t *p, uint4 v) { stg_cs_v4(p, v); }
};
template <> struct VecIO<uint2> {
    static __device__ __forceinline__ uint2 ld(const uint8_t *p) { return __ldg(reinterpret_cast<const uint2 *>(p)); }
    static __device__ __forceinline__ void st(uint8_t *p, uint2 v) { stg_cs_v2(p, v); }
};
template <> struct VecIO<uint32_t> {
    static __device__ __forceinline__ uint32_t ld(const uint8_t *p) { return __ldg(reinterpret_cast<const uint32_t *>(p)); }
    static __device__ __forceinline__ void st(uint8_t *p, uint32_t v) { stg_cs_u32(p, v); }
};

__device__ __forceinline__ bool item_to_entry(int32_t it, uint32_t T, uint32_t &entry) {
    const uint32_t a = (uint32_t)(it < 0 ? -it : it);
    if (it == 0 || a > T) return false;
    entry = 2 * (a - 1) + (it < 0 ? 1u : 0u);
    return true;
}

// One thread owns one V-wide column piece of one block row and walks the ts tile rows:
// consecutive threads -> consecutive bytes of the output row (fully coalesced V-wide stores),
// the ts independent loads of a thread give the memory-level parallelism.
template <typename V>
__global__ void __launch_bounds__(256) compose_copy_kernel(const uint8_t *__restrict__ lib, const int32_t *__restrict__ item,
                                                           uint32_t T, uint32_t ts, uint32_t bw, uint8_t *__restrict__ out,
                                                           int *__restrict__ err) {
    const uint32_t RB = ts * 3;                       // bytes per tile row
    const uint32_t pieces = bw * RB / sizeof(V);      // per output row
    const uint32_t piece = blockIdx.x * blockDim.x + threadIdx.x;
    if (piece >= pieces) return;
    const uint32_t by = blockIdx.y;
    const uint32_t byte0 = piece * (uint32_t)sizeof(V);
    const uint32_t bx = byte0 / RB, off = byte0 % RB;
    uint32_t entry;
    if (!item_to_entry(item[(size_t)by * bw + bx], T, entry)) {
        atomicOr(err, 1);
        return;
    }
    const uint8_t *s = lib + (size_t)entry * ts * RB + off;
    const size_t OWB = (size_t)bw * RB;
    uint8_t *d = out + (size_t)by * ts * OWB + byte0;
    uint32_t r = 0;
    for (; r + 8 <= ts; r += 8) {
        V v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = VecIO<V>::ld(s + (size_t)(r + k) * RB);
#pragma unroll
        for (int k = 0; k < 8; k++) VecIO<V>::st(d + (size_t)(r + k) * OWB, v[k]);
    }
    for (; r < ts; r++) VecIO<V>::st(d + (size_t)r * OWB, VecIO<V>::ld(s + (size_t)r * RB));
}

// Small tiles (ts = 8, 16): a tile row is only 24 / 48 bytes, so the row-wise gather above costs one
// L1 wavefront per 24..48 bytes and the kernel becomes L1TEX-bound (ncu: l1tex 96 %, dram 60 %).
// Here a CTA takes G consecutive tiles of one block row (G*3*ts = 1536 B of output per row), reads each
// tile as ONE contiguous run (192 / 768 B, 16-byte pieces), transposes through shared memory and writes
// every output row chunk as 1536 contiguous bytes with 16-byte stores.
template <int TS>
__global__ void __launch_bounds__(192) compose_tile_kernel(const uint8_t *__restrict__ lib, const int32_t *__restrict__ item,
                                                           uint32_t T, uint32_t bw, uint8_t *__restrict__ out,
                                                           int *__restrict__ err) {
    constexpr int RB = TS * 3;                 // bytes per tile row
    constexpr int TILE_B = TS * RB;            // bytes per tile
    constexpr int G = 1536 / RB;               // tiles per CTA
    constexpr int ROW_CHUNK = G * RB;          // 1536
    constexpr int STRIDE = ROW_CHUNK + 16;     // padded shared-memory row stride (bank spread)
    constexpr int PPT = TILE_B / 16;           // 16-byte pieces per tile (12 / 48)
    constexpr int NTHR = 192;                  // a multiple of PPT: a thread keeps its piece index in every pass
    constexpr int TPP = NTHR / PPT;            // tiles per pass (16 / 4)
    constexpr int PASSES = G / TPP;            // 4 / 8
    __shared__ __align__(16) uint8_t sm[TS * STRIDE];
    const uint32_t by = blockIdx.y, bx0 = blockIdx.x * G;
    const int part = threadIdx.x % PPT, tile0 = threadIdx.x / PPT;
    const int32_t *it = item + (size_t)by * bw + bx0 + tile0;
    grid_dependency_wait();  // the item map comes from the kernel before this one
    uint4 v[PASSES];
#pragma unroll
    for (int k = 0; k < PASSES; k++) {
        uint32_t e = 0;
        if (!item_to_entry(__ldg(it + k * TPP), T, e)) atomicOr(err, 1);
        v[k] = __ldg(reinterpret_cast<const uint4 *>(lib + (size_t)e * TILE_B + part * 16));
    }
    // shared-memory destination(s) of this thread's piece inside a tile: constant over the passes
    uint8_t *dst = sm + tile0 * RB;
    if constexpr (RB % 16 == 0) {
        dst += ((part * 16) / RB) * STRIDE + (part * 16) % RB;
#pragma unroll
        for (int k = 0; k < PASSES; k++) *reinterpret_cast<uint4 *>(dst + k * TPP * RB) = v[k];
    } else {  // RB = 24: the piece straddles rows at 8-byte granularity
        const int h0 = 2 * part, h1 = 2 * part + 1;
        uint8_t *d0 = dst + (h0 / 3) * STRIDE + (h0 % 3) * 8, *d1 = dst + (h1 / 3) * STRIDE + (h1 % 3) * 8;
#pragma unroll
        for (int k = 0; k < PASSES; k++) {
            *reinterpret_cast<uint2 *>(d0 + k * TPP * RB) = make_uint2(v[k].x, v[k].y);
            *reinterpret_cast<uint2 *>(d1 + k * TPP * RB) = make_uint2(v[k].z, v[k].w);
        }
    }
    // make the generic-proxy shared-memory writes visible to the async proxy, then let the TMA engine
    // stream every output row chunk (1536 contiguous bytes) to HBM: no LDS/STG on the LSU data pipe
    fence_proxy_async_smem();
    __syncthreads();
    const size_t OWB = (size_t)bw * RB;
    uint8_t *d = out + (size_t)by * TS * OWB + (size_t)bx0 * RB;
    if (threadIdx.x < TS) {
        // the output is written once and never read on the device: first in line for eviction, so that the tile
        // library, the item map and the search index keep their place in L2 (C4 compose: 514 -> 500 us)
        bulk_s2g_hint(d + (size_t)threadIdx.x * OWB, sm + threadIdx.x * STRIDE, ROW_CHUNK, l2_policy_evict_first());
        bulk_commit_wait_read();  // shared memory must stay alive until the engine has read it
    }
}

// ---------------------------------------------------------------------------------------
// generic kernel: one thread per output pixel, any geometry, RGB or RGBA(+tint via exact table)
// ---------------------------------------------------------------------------------------
__global__ void compose_generic_kernel(const uint8_t *__restrict__ lib, const int32_t *__restrict__ item, uint32_t T,
                                       uint32_t ts, uint32_t dim, uint32_t bw, uint32_t bh, const uint8_t *__restrict__ src,
                                       uint32_t W, uint32_t H, uint32_t oc, const uint8_t *__restrict__ lut, uint8_t alpha_out,
                                       uint8_t *__restrict__ out, int *__restrict__ err) {
    const uint64_t OW = (uint64_t)bw * ts, OH = (uint64_t)bh * ts, total = OW * OH;
    const float xr = (float)W / (float)OW, yr = (float)H / (float)OH;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t Y = (uint32_t)(i / OW), X = (uint32_t)(i % OW);
        const uint32_t by = Y / ts, r = Y % ts, bx = X / ts, c = X % ts;
        uint32_t entry;
        if (!item_to_entry(item[(size_t)by * bw + bx], T, entry)) {
            atomicOr(err, 1);
            continue;
        }
        const uint8_t *p = lib + ((size_t)entry * ts * ts + (size_t)r * ts + c) * 3;
        uint8_t v0 = p[0], v1 = p[1], v2 = p[2];
        if (oc == 3) {
            uint8_t *o = out + i * 3;
            o[0] = v0; o[1] = v1; o[2] = v2;
        } else {
            // image 0.25.2 resize(Nearest): source pixel floor((X + 0.5) * W/OW), in f32
            uint32_t sx = (uint32_t)floorf(__fmul_rn(__fadd_rn((float)X, 0.5f), xr));
            uint32_t sy = (uint32_t)floorf(__fmul_rn(__fadd_rn((float)Y, 0.5f), yr));
            sx = sx > W - 1 ? W - 1 : sx;
            sy = sy > H - 1 ? H - 1 : sy;
            const uint8_t *f = src + ((size_t)sy * W + sx) * 3;
            v0 = lut[v0 * 256 + f[0]];
            v1 = lut[v1 * 256 + f[1]];
            v2 = lut[v2 * 256 + f[2]];
            reinterpret_cast<uint32_t *>(out)[i] = v0 | (v1 << 8) | (v2 << 16) | ((uint32_t)alpha_out << 24);
        }
    }
}

// ---------------------------------------------------------------------------------------
// fast tint kernel: 4 pixels (12 B in, 16 B out) per thread per row, K exceptions per channel
// ---------------------------------------------------------------------------------------
template <int K>
__device__ __forceinline__ uint32_t blend_px(uint32_t p, const uint32_t (&ew)[K > 0 ? K : 1], const uint32_t (&mw)[K > 0 ? K : 1],
                                             uint32_t kmul, uint32_t c_rb, uint32_t c_ga) {
    // p = bg pixel (r, g, b, junk). Exceptional bg byte -> bg-1 (gives q-1 exactly, see header).
#pragma unroll
    for (int k = 0; k < K; k++) {
        const uint32_t t = p ^ ew[k];
        const uint32_t u = (t & 0x7f7f7f7fu) + 0x7f7f7f7fu;
        const uint32_t z = ~(u | t) & mw[k];  // 0x80 in every enabled byte where bg == exceptional value
        p -= z >> 7;
    }
    const uint32_t rb = p & 0x00ff00ffu;
    const uint32_t g = __byte_perm(p, 0, 0x4441);          // (g, 0, 0, 0)
    const uint32_t xrb = rb * kmul + c_rb;                  // two 16-bit lanes: bg*(255-A) + fg*A + 1
    const uint32_t xga = g * kmul + c_ga;                   // lane 1 carries the constant alpha byte
    const uint32_t zrb = xrb + __byte_perm(xrb, 0, 0x4341); // x + (x >> 8) per lane
    const uint32_t zga = xga + __byte_perm(xga, 0, 0x4341);
    return __byte_perm(zrb, zga, 0x7351);                   // (r, g, b, a) = high bytes of the lanes
}

template <int K>
__global__ void __launch_bounds__(256) compose_tint_kernel(const uint8_t *__restrict__ lib, const int32_t *__restrict__ item,
                                                           uint32_t T, uint32_t ts, uint32_t dim, uint32_t bw,
                                                           const uint8_t *__restrict__ src, uint32_t W, uint32_t A,
                                                           uint32_t alpha_out, const uint8_t *__restrict__ excv,
                                                           const uint8_t *__restrict__ excm,
                                                           const uint8_t *__restrict__ cadd, uint8_t *__restrict__ out,
                                                           int *__restrict__ err) {
    const uint32_t gpr = ts / 4;                      // 4-pixel groups per tile row
    const uint32_t groups = bw * gpr;                 // per output row
    const uint32_t grp = blockIdx.x * blockDim.x + threadIdx.x;
    if (grp >= groups) return;
    const uint32_t by = blockIdx.y;
    const uint32_t bx = grp / gpr, g4 = grp % gpr;
    uint32_t entry;
    if (!item_to_entry(item[(size_t)by * bw + bx], T, entry)) {
        atomicOr(err, 1);
        return;
    }
    const uint32_t RB = ts * 3, cell = ts / dim;      // cell: output pixels per source pixel
    const uint8_t *s = lib + (size_t)entry * ts * RB + g4 * 12;
    const size_t OWB = (size_t)bw * ts * 4;
    uint8_t *d = out + (size_t)by * ts * OWB + (size_t)grp * 16;
    const uint32_t kmul = 255 - A;
    const uint32_t sx = bx * dim + (g4 * 4) / cell;
    for (uint32_t cr = 0; cr < dim; cr++) {
        const uint8_t *f = src + ((size_t)(by * dim + cr) * W + sx) * 3;
        const uint32_t fr = f[0], fg = f[1], fb = f[2];
        const uint32_t c_rb = (fr * A + cadd[fr]) | ((fb * A + cadd[fb]) << 16);
        const uint32_t c_ga = (fg * A + cadd[fg]) | (alpha_out << 24);  // lane 1 = alpha_out * 256
        uint32_t ew[K > 0 ? K : 1], mw[K > 0 ? K : 1];
#pragma unroll
        for (int k = 0; k < K; k++) {
            ew[k] = excv[k * 256 + fr] | (excv[k * 256 + fg] << 8) | (excv[k * 256 + fb] << 16);
            mw[k] = excm[k * 256 + fr] | (excm[k * 256 + fg] << 8) | (excm[k * 256 + fb] << 16);
        }
#pragma unroll 4
        for (uint32_t r = cr * cell; r < (cr + 1) * cell; r++) {
            const uint32_t *in = reinterpret_cast<const uint32_t *>(s + (size_t)r * RB);
            const uint32_t w0 = __ldg(in), w1 = __ldg(in + 1), w2 = __ldg(in + 2);
            uint4 o;
            o.x = blend_px<K>(w0, ew, mw, kmul, c_rb, c_ga);
            o.y = blend_px<K>(__byte_perm(w0, w1, 0x6543), ew, mw, kmul, c_rb, c_ga);
            o.z = blend_px<K>(__byte_perm(w1, w2, 0x5432), ew, mw, kmul, c_rb, c_ga);
            o.w = blend_px<K>(w2 >> 8, ew, mw, kmul, c_rb, c_ga);
            stg_cs_v4(d + (size_t)r * OWB, o);
        }
    }
}

int emo_launch_compose(emo_ctx *ctx, const int32_t *item, const uint8_t *src, uint32_t W, uint32_t H, uint32_t oc,
                       uint8_t tint_alpha, uint8_t *out) {
    const uint32_t ts = ctx->ts, dim = ctx->dim, bw = W / dim, bh = H / dim, T = ctx->T;
    const uint32_t RB = ts * 3;
    const bool aligned = ((uintptr_t)out % 16 == 0) && bh <= 65535;
    static const bool pdl = !(getenv("EMO_PDL") && atoi(getenv("EMO_PDL")) == 0);  // EMO_PDL=0: plain launches
    if (oc == 3) {
        if (aligned && ts == 8 && bw % 64 == 0) {
            if (pdl) EMO_CK(emo_launch_pdl(compose_tile_kernel<8>, dim3(bw / 64, bh), dim3(192), 0, ctx->stream, ctx->lib_px, item, T, bw, out, ctx->err_flag));
            else compose_tile_kernel<8><<<dim3(bw / 64, bh), 192, 0, ctx->stream>>>(ctx->lib_px, item, T, bw, out, ctx->err_flag);
        } else if (aligned && ts == 16 && bw % 32 == 0) {
            if (pdl) EMO_CK(emo_launch_pdl(compose_tile_kernel<16>, dim3(bw / 32, bh), dim3(192), 0, ctx->stream, ctx->lib_px, item, T, bw, out, ctx->err_flag));
            else compose_tile_kernel<16><<<dim3(bw / 32, bh), 192, 0, ctx->stream>>>(ctx->lib_px, item, T, bw, out, ctx->err_flag);
        } else if (aligned && RB % 16 == 0) {
            const uint32_t pieces = bw * RB / 16;
            compose_copy_kernel<uint4><<<dim3((pieces + 255) / 256, bh), 256, 0, ctx->stream>>>(ctx->lib_px, item, T, ts, bw, out,
                                                                                                ctx->err_flag);
        } else if (aligned && RB % 8 == 0) {
            const uint32_t pieces = bw * RB / 8;
            compose_copy_kernel<uint2><<<dim3((pieces + 255) / 256, bh), 256, 0, ctx->stream>>>(ctx->lib_px, item, T, ts, bw, out,
                                                                                                ctx->err_flag);
        } else if (aligned && RB % 4 == 0) {
            const uint32_t pieces = bw * RB / 4;
            compose_copy_kernel<uint32_t><<<dim3((pieces + 255) / 256, bh), 256, 0, ctx->stream>>>(ctx->lib_px, item, T, ts, bw,
                                                                                                   out, ctx->err_flag);
        } else {
            const uint64_t total = (uint64_t)bw * ts * bh * ts;
            const uint64_t blocks = (total + 255) / 256, cap = (uint64_t)ctx->sm_count * 32;
            compose_generic_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, ctx->stream>>>(
                ctx->lib_px, item, T, ts, dim, bw, bh, nullptr, W, H, 3, nullptr, 255, out, ctx->err_flag);
        }
        EMO_LAUNCH_CHECK(ctx);
        return EMO_OK;
    }
    int rc = emo_prepare_tint(ctx, tint_alpha);
    if (rc) return rc;
    const emo_tint_tables &t = ctx->tint;
    const uint32_t cell = ts / dim;
    if (aligned && t.K <= 3 && ts % 4 == 0 && cell % 4 == 0) {
        const uint32_t groups = bw * (ts / 4);
        dim3 grid((groups + 255) / 256, bh);
#define EMO_TINT(KK)                                                                                                   \
    compose_tint_kernel<KK><<<grid, 256, 0, ctx->stream>>>(ctx->lib_px, item, T, ts, dim, bw, src, W, tint_alpha,       \
                                                           t.alpha_out, t.excv, t.excm, t.cadd, out, ctx->err_flag)
        switch (t.K) {
            case 0: EMO_TINT(0); break;
            case 1: EMO_TINT(1); break;
            case 2: EMO_TINT(2); break;
            default: EMO_TINT(3); break;
        }
#undef EMO_TINT
    } else {
        const uint64_t total = (uint64_t)bw * ts * bh * ts;
        const uint64_t blocks = (total + 255) / 256, cap = (uint64_t)ctx->sm_count * 32;
        compose_generic_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, ctx->stream>>>(
            ctx->lib_px, item, T, ts, dim, bw, bh, src, W, H, 4, t.lut, t.alpha_out, out, ctx->err_flag);
    }
    EMO_LAUNCH_CHECK(ctx);
    return EMO_OK;
}

// Tint with an overlay image whose size differs from the matched source (src/main.rs:447-478 overlays the ORIGINAL
// image, which n_to_1 may have resized before matching: --downsample > 1 or dimensions not divisible by dim).
// Any ratio: the generic kernel samples floor((X + 0.5) * ow / OW) like image 0.25.2 resize(Nearest).
int emo_launch_compose_overlay(emo_ctx *ctx, const int32_t *item, uint32_t W, uint32_t H, const uint8_t *overlay, uint32_t ow,
                               uint32_t oh, uint8_t tint_alpha, uint8_t *out) {
    const uint32_t ts = ctx->ts, dim = ctx->dim, bw = W / dim, bh = H / dim;
    int rc = emo_prepare_tint(ctx, tint_alpha);
    if (rc) return rc;
    const emo_tint_tables &t = ctx->tint;
    const uint64_t total = (uint64_t)bw * ts * bh * ts;
    const uint64_t blocks = (total + 255) / 256, cap = (uint64_t)ctx->sm_count * 32;
    compose_generic_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, ctx->stream>>>(
        ctx->lib_px, item, ctx->T, ts, dim, bw, bh, overlay, ow, oh, 4, t.lut, t.alpha_out, out, ctx->err_flag);
    EMO_LAUNCH_CHECK(ctx);
    return EMO_OK;
}
