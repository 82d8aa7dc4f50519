// resize.cu — image 0.25.2 imageops::resize(view, nw, nh, FilterType::Lanczos3), bit-exact, for RGB8 batches.
//
// Reference call sites (paths under the reference root): src/main.rs:595 (the source image is resized to dimensions
// divisible by the cell grid / by --downsample before matching) and src/mosaic/tiles/utils.rs:188-189 (prepare_tile
// resizes the trimmed / centre-cropped view of a photo to tile_size x tile_size).  The crate's algorithm
// (imageops/sample.rs) is two separable passes with an f32 intermediate image:
//   vertical_sample:   tmp[oy][x][c] = sum_i f32(px[left(oy) + i][x][c]) * w(oy, i)        (kept in f32)
//   horizontal_sample: out[y][ox][c] = round(clamp(sum_i tmp[y][left(ox) + i][c] * w(ox, i), 0, 255))
// with every tap one f32 multiply followed by one f32 add (Rust never contracts to FMA), taps in ascending order.
// f32 addition is not associative, so a tap loop is inherently sequential per output element and both kernels keep
// that order; the parallelism is across output elements.  Every rounding of the reference is kept: the horizontal pass
// issues the multiply and the add as __fmul_rn / __fadd_rn so ptxas cannot fuse them, the vertical pass forms the
// once-rounded product with an FMA (below) and adds separately.
//
// The tap tables (left, count, normalised Lanczos3 weights: O(nw + nh) numbers, each needing the platform's libm sinf
// exactly as f32::sin does in the reference) are computed on the host in resize_axis() and cached per geometry; all
// O(pixels) arithmetic runs here.
//
//   resize_vertical2_kernel<TRANSPOSED, DEPTH, NB>: the vertical pass of every 8-byte aligned geometry — two output rows per
//     thread share the loads and the byte -> 2^23 + b step where their windows overlap (see the kernel).
//   resize_vertical_kernel<ALIGNED, TRANSPOSED>: one thread per (image, output row, 8 consecutive bytes of the view row); the
//     path of rows that are not 8-byte aligned and of tap tables whose neighbouring windows are not monotone.
//     Rows of the view are read as one 64-bit word when the geometry keeps them 8-byte aligned, else as the two or three
//     covering aligned 32-bit words funnel-shifted by the row's phase.  The product f32(b) * w of a tap is formed WITHOUT
//     converting the byte first: PRMT builds x = 0x4B0000bb = 2^23 + b (exact), and fma(x, w, -2^23 * w) =
//     RN((2^23 + b) * w - 2^23 * w) = RN(b * w), because the FMA adds the exact 48-bit product to the exactly representable
//     -2^23 * w and rounds once — the same single rounding as the reference's multiply.  The add that accumulates stays
//     a separate FADD: 3 instructions per byte and tap (PRMT, FFMA, FADD) instead of 4.  w and -2^23 * w come from the
//     table as one 128-bit load per four taps each.  The kernel is issue-bound (ncu, 64 x 2048^2 -> 64^2: 70 % of the issue
//     slots, ALU pipe 53 %, DRAM 14 %, L2 32 %; DRAM reads = the algorithmic bytes): the ~6x re-read of a source row by
//     neighbouring output rows is served by L1/L2 and is not the limiter.
//   tmp layout: row-major [oy][x*3+c] (pitch padded to 8 floats, two float4 stores per thread) when the image is about as
//     wide as the output (source resize); TRANSPOSED [x*3+c][oy] for batches whose axes both shrink 8x or more (photo -> tile): there
//     the windows of neighbouring output pixels lie far apart, so the horizontal pass runs with lanes along oy and reads
//     every tap as one coalesced 128-byte row of the transposed image (the row-major layout costs 32 cache lines per load).
//   resize_horizontal_kernel: one thread per (image, row, output pixel), lanes along the output row, weights stored
//     tap-major ([tap][ox]) so that lanes read consecutive words.
//   resize_horizontal_smem_kernel<WS>: row-major input whose output pixels are 4+ source pixels apart (a single photo,
//     batches of very small tiles): the tmp row is staged into shared memory with coalesced loads (skewed by one word per 32
//     so that lanes 3*ratio floats apart hit different banks) and each thread runs its tap chain from there — the global
//     version of this access pattern costs 32 cache lines per load.  WS: fewer blocks than SMs, the weights are staged too.
//   resize_horizontal_t_kernel: transposed input, one block per (image, output column), lanes along oy, weights uniform.
//   resize_copy_kernel: the "(nwidth, nheight) == image.dimensions()" early return of resize() — a plain copy.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

// ---- host: tap tables (sample.rs horizontal_sample / vertical_sample set-up) -------------------------------------------
static float lanczos3_sinc(float t) {
    const float a = t * 3.14159274101257324f;  // f32::consts::PI
    return t == 0.0f ? 1.0f : sinf(a) / a;
}
static float lanczos3(float x) { return fabsf(x) < 3.0f ? lanczos3_sinc(x) * lanczos3_sinc(x / 3.0f) : 0.0f; }

struct emo_resize_axis {
    uint32_t n_in = 0, n_out = 0, pitch = 0;  // pitch = largest tap count
    std::vector<uint32_t> left, cnt;
    std::vector<float> ws;  // [n_out][pitch] (row-major per output index), zero beyond cnt
};

// volatile stores keep every intermediate an IEEE f32 whatever the host compiler's contraction setting is
static void resize_axis(uint32_t n_in, uint32_t n_out, emo_resize_axis &ax) {
    ax.n_in = n_in;
    ax.n_out = n_out;
    ax.left.assign(n_out, 0);
    ax.cnt.assign(n_out, 0);
    const float ratio = (float)n_in / (float)n_out;
    const float sratio = ratio < 1.0f ? 1.0f : ratio;
    const float support = 3.0f * sratio;
    uint32_t pitch = 0;
    std::vector<float> centre(n_out);
    for (uint32_t o = 0; o < n_out; o++) {
        volatile float c = ((float)o + 0.5f) * ratio;
        volatile float lo = c - support, hi = c + support;
        long long l = (long long)floorf(lo);
        if (l < 0) l = 0;
        if (l > (long long)n_in - 1) l = (long long)n_in - 1;
        long long r = (long long)ceilf(hi);
        if (r < l + 1) r = l + 1;
        if (r > (long long)n_in) r = (long long)n_in;
        volatile float c0 = c - 0.5f;
        centre[o] = c0;
        ax.left[o] = (uint32_t)l;
        ax.cnt[o] = (uint32_t)(r - l);
        if (ax.cnt[o] > pitch) pitch = ax.cnt[o];
    }
    ax.pitch = pitch;
    ax.ws.assign((size_t)n_out * pitch, 0.0f);
    for (uint32_t o = 0; o < n_out; o++) {
        float *w = ax.ws.data() + (size_t)o * pitch;
        volatile float sum = 0.0f;
        for (uint32_t i = 0; i < ax.cnt[o]; i++) {
            volatile float d = (float)(ax.left[o] + i) - centre[o];
            volatile float x = d / sratio;
            w[i] = lanczos3(x);
            sum = sum + w[i];
        }
        for (uint32_t i = 0; i < ax.cnt[o]; i++) w[i] = w[i] / sum;
    }
}

// Diagnostic export: the tap table of one axis exactly as emo_resize computes it (host-only, no ctx, no GPU).
extern "C" int emo_resize_taps(uint32_t n_in, uint32_t n_out, uint32_t *left, uint32_t *cnt, float *ws, uint32_t pitch,
                               uint32_t *max_taps) {
    EMO_REQUIRE(n_in >= 1 && n_out >= 1, EMO_ERR_ARG, "resize_taps: empty axis (%u -> %u)", n_in, n_out);
    emo_resize_axis ax;
    resize_axis(n_in, n_out, ax);
    if (max_taps) *max_taps = ax.pitch;
    if (left) memcpy(left, ax.left.data(), (size_t)n_out * 4);
    if (cnt) memcpy(cnt, ax.cnt.data(), (size_t)n_out * 4);
    if (ws) {
        EMO_REQUIRE(pitch >= ax.pitch, EMO_ERR_ARG, "resize_taps: pitch %u below the largest tap count %u", pitch, ax.pitch);
        for (uint32_t o = 0; o < n_out; o++) {
            memset(ws + (size_t)o * pitch, 0, (size_t)pitch * 4);
            memcpy(ws + (size_t)o * pitch, ax.ws.data() + (size_t)o * ax.pitch, (size_t)ax.cnt[o] * 4);
        }
    }
    return EMO_OK;
}

// ---- device ---------------------------------------------------------------------------------------------------------------
// RN(f32(byte K of v) * w) in one FFMA: x = 2^23 + b (PRMT), nbw = -2^23 * w (exact), fma(x, w, nbw) rounds b * w once
template <int K>
__device__ __forceinline__ float byte_times(uint32_t v, float w, float nbw) {
    return __fmaf_rn(__uint_as_float(__byte_perm(v, 0x4B000000u, 0x7540 | K)), w, nbw);
}

// one tap of one 4-byte group: t[k] += RN(byte_k * w), the multiply as the exact FFMA above, the add separate
__device__ __forceinline__ void tap4(float *t, uint32_t v, float w, float nbw) {
    t[0] = __fadd_rn(t[0], byte_times<0>(v, w, nbw));
    t[1] = __fadd_rn(t[1], byte_times<1>(v, w, nbw));
    t[2] = __fadd_rn(t[2], byte_times<2>(v, w, nbw));
    t[3] = __fadd_rn(t[3], byte_times<3>(v, w, nbw));
}

template <bool ALIGNED, bool TRANSPOSED>
__global__ void __launch_bounds__(256)
resize_vertical_kernel(const uint8_t *__restrict__ src, size_t img_bytes, uint32_t row_stride, size_t base_off, uint32_t row_bytes,
                       const uint32_t *__restrict__ left, const uint32_t *__restrict__ cnt, const float *__restrict__ ws,
                       const float *__restrict__ nbws, uint32_t wpitch, float *__restrict__ tmp, uint32_t tpitch, uint32_t nh,
                       uint32_t np, uint32_t n0) {
    const uint32_t xb = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (xb >= row_bytes) return;
    const uint32_t oy = blockIdx.y, n = blockIdx.z + n0;
    const uint32_t c = cnt[oy];
    const float *__restrict__ w = ws + (size_t)oy * wpitch;      // wpitch is a multiple of 4: 16-byte aligned rows
    const float *__restrict__ nb = nbws + (size_t)oy * wpitch;   // -2^23 * w
    const uint8_t *p = src + (size_t)n * img_bytes + base_off + (size_t)left[oy] * row_stride + xb;
    const uint32_t valid = row_bytes - xb;  // bytes of this thread's group that belong to the view (>= 1)
    float t[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    if (ALIGNED && valid >= 8) {
        uint32_t i = 0;
        for (; i + 4 <= c; i += 4) {
            const float4 w4 = __ldg((const float4 *)(w + i)), n4 = __ldg((const float4 *)(nb + i));
            const uint2 v0 = __ldg((const uint2 *)p), v1 = __ldg((const uint2 *)(p + row_stride)),
                        v2 = __ldg((const uint2 *)(p + 2 * (size_t)row_stride)), v3 = __ldg((const uint2 *)(p + 3 * (size_t)row_stride));
            p += 4 * (size_t)row_stride;
            tap4(t, v0.x, w4.x, n4.x); tap4(t + 4, v0.y, w4.x, n4.x);
            tap4(t, v1.x, w4.y, n4.y); tap4(t + 4, v1.y, w4.y, n4.y);
            tap4(t, v2.x, w4.z, n4.z); tap4(t + 4, v2.y, w4.z, n4.z);
            tap4(t, v3.x, w4.w, n4.w); tap4(t + 4, v3.y, w4.w, n4.w);
        }
        for (; i < c; i++, p += row_stride) {
            const uint2 v = __ldg((const uint2 *)p);
            const float wi = __ldg(w + i), ni = __ldg(nb + i);
            tap4(t, v.x, wi, ni); tap4(t + 4, v.y, wi, ni);
        }
    } else {
        // rows that are not 8-byte aligned (the usual case when the width is not divisible: 3 * 4097 bytes per row): the
        // aligned 32-bit words that cover the group, funnel-shifted by the row's phase.  An aligned word never straddles a
        // page, and a word is only touched when it holds a byte of the view.
        const uint32_t need = valid < 8 ? valid : 8;
#pragma unroll 2
        for (uint32_t i = 0; i < c; i++, p += row_stride) {
            const uint32_t ph = (uint32_t)((uintptr_t)p & 3);
            const uint32_t *q = (const uint32_t *)(p - ph);
            const uint32_t a0 = __ldg(q);
            const uint32_t a1 = ph + need > 4 ? __ldg(q + 1) : 0u;
            const uint32_t a2 = ph + need > 8 ? __ldg(q + 2) : 0u;
            const float wi = __ldg(w + i), ni = __ldg(nb + i);
            // bytes beyond the view only feed the padding lanes of tmp
            tap4(t, __funnelshift_r(a0, a1, ph * 8), wi, ni);
            tap4(t + 4, __funnelshift_r(a1, a2, ph * 8), wi, ni);
        }
    }
    if (TRANSPOSED) {
        // [x*3+c][oy], np floats per line; the padding lines of the last group (sums of bytes beyond the view) exist in the buffer
        float *o = tmp + ((size_t)blockIdx.z * tpitch + xb) * np + oy;
#pragma unroll
        for (int k = 0; k < 8; k++) o[k * (size_t)np] = t[k];
    } else {
        // rows are padded to a multiple of 8 floats: the padding lanes of the last group are never read
        float4 *o = (float4 *)(tmp + ((size_t)blockIdx.z * nh + oy) * tpitch + xb);
        o[0] = make_float4(t[0], t[1], t[2], t[3]);
        o[1] = make_float4(t[4], t[5], t[6], t[7]);
    }
}

// Two output rows per thread (aligned geometries).  The windows of neighbouring output rows overlap — by 5/6 of their length when
// the image shrinks — so a thread that owns rows oy and oy + 1 of its NB-byte column loads every source row once and builds
// 2^23 + b once for both: per byte and tap PRMT / 2 + FFMA + FADD = 2.5 issue slots instead of 3.  The block first lays the two
// weight rows side by side in shared memory, indexed by source row: {wA, -2^23 wA, wB, -2^23 wB}, zero where a row lies outside a
// window (a zero tap adds +0 to an accumulator that is never -0: exact), so the loop has no per-row tests; whole groups of four
// source rows that touch one window only run a single-output body (block-uniform branch).  Source rows are loaded DEPTH groups
// ahead of their use into register sets that rotate by unrolling.  NB = 8 bytes per thread, or 4 when the grid would not fill
// the GPU otherwise (a single photo: twice the warps).
template <bool TRANSPOSED, int DEPTH, int NB>
__global__ void __launch_bounds__(256)
resize_vertical2_kernel(const uint8_t *__restrict__ src, size_t img_bytes, uint32_t row_stride, size_t base_off, uint32_t row_bytes,
                        const uint32_t *__restrict__ left, const uint32_t *__restrict__ cnt, const float *__restrict__ ws,
                        uint32_t wpitch, float *__restrict__ tmp, uint32_t tpitch, uint32_t nh, uint32_t np, uint32_t n0) {
    constexpr int NW = NB / 4;  // 32-bit words per row and thread
    extern __shared__ float4 tab[];  // [4 * ng]
    const uint32_t oyA = 2 * blockIdx.y, oyB = oyA + 1, n = blockIdx.z + n0;
    const bool hasB = oyB < nh;
    const uint32_t lA = left[oyA], eA = lA + cnt[oyA];
    const uint32_t lB = hasB ? left[oyB] : eA, eB = hasB ? lB + cnt[oyB] : eA;  // lA <= lB and eA <= eB (checked on the host)
    const uint32_t span = eB - lA, ng = (span + 3) / 4;
    for (uint32_t j = threadIdx.x; j < 4 * ng; j += blockDim.x) {
        const uint32_t r = lA + j;
        const float wA = r < eA ? __ldg(ws + (size_t)oyA * wpitch + j) : 0.0f;
        const float wB = (r >= lB && r < eB) ? __ldg(ws + (size_t)oyB * wpitch + (r - lB)) : 0.0f;
        tab[j] = make_float4(wA, wA * -8388608.0f, wB, wB * -8388608.0f);  // a power of two: exact
    }
    __syncthreads();
    const uint32_t xb = (blockIdx.x * blockDim.x + threadIdx.x) * NB;
    if (xb >= row_bytes) return;
    // the aligned 8-byte word that holds a byte of the view lies inside the image row (row starts and strides are multiples of 8)
    const uint8_t *pl = src + (size_t)n * img_bytes + base_off + (size_t)lA * row_stride + xb;  // first row of the next group to load
    uint32_t gl = 0;
    const uint32_t last = span - 1;
    const uint32_t gA_end = (eA - lA + 3) / 4;         // groups below touch window A
    const uint32_t gB_beg = hasB ? (lB - lA) / 4 : ng;  // groups from here on touch window B
    float tA[NB], tB[NB];
#pragma unroll
    for (int k = 0; k < NB; k++) tA[k] = tB[k] = 0.0f;
    auto load_row = [&](uint32_t (&w)[NW], const uint8_t *q) {
        if (NW == 2) {
            const uint2 t = __ldg((const uint2 *)q);
            w[0] = t.x; w[NW - 1] = t.y;
        } else {
            w[0] = __ldg((const uint32_t *)q);
        }
    };
    // the next group's rows into v; only the last group can reach beyond the span (its extra rows have zero weights: any row will do)
    auto load_next = [&](uint32_t (&v)[4][NW]) {
        if (gl + 1 < ng) {
#pragma unroll
            for (uint32_t k = 0; k < 4; k++) load_row(v[k], pl + (size_t)k * row_stride);
            pl += 4 * (size_t)row_stride;
        } else if (gl < ng) {
            const uint32_t left_rows = last - 4 * gl;  // >= 0: the group holds at least one row of the span
#pragma unroll
            for (uint32_t k = 0; k < 4; k++) load_row(v[k], pl + (size_t)min(k, left_rows) * row_stride);
        }
        gl++;
    };
    auto taps = [&](const uint32_t (&v)[4][NW], uint32_t g) {
        const float4 *tg = tab + 4 * g;
        const bool a = g < gA_end, b = g >= gB_beg;
        if (a && b) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float4 w = tg[k];
#pragma unroll
                for (int h = 0; h < NW; h++) {
                    tap4(tA + 4 * h, v[k][h], w.x, w.y);
                    tap4(tB + 4 * h, v[k][h], w.z, w.w);
                }
            }
        } else if (a) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float2 w = *(const float2 *)(tg + k);
#pragma unroll
                for (int h = 0; h < NW; h++) tap4(tA + 4 * h, v[k][h], w.x, w.y);
            }
        } else if (b) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float2 w = *((const float2 *)(tg + k) + 1);
#pragma unroll
                for (int h = 0; h < NW; h++) tap4(tB + 4 * h, v[k][h], w.x, w.y);
            }
        }
    };
    uint32_t v[DEPTH + 1][4][NW];
#pragma unroll
    for (int s = 0; s < DEPTH; s++) load_next(v[s]);
    for (uint32_t g = 0; g < ng; g += DEPTH + 1) {
#pragma unroll
        for (int s = 0; s <= DEPTH; s++) {
            load_next(v[(s + DEPTH) % (DEPTH + 1)]);
            if (g + s < ng) taps(v[s], g + s);
        }
    }
    if (TRANSPOSED) {
        // [x*3+c][oy], np floats per line; the padding lines of the last group exist in the buffer
        float *o = tmp + ((size_t)blockIdx.z * tpitch + xb) * np + oyA;
#pragma unroll
        for (int k = 0; k < NB; k++) {
            if (hasB) *(float2 *)(o + k * (size_t)np) = make_float2(tA[k], tB[k]);  // oyA is even, np a multiple of 32
            else o[k * (size_t)np] = tA[k];
        }
    } else {
        // rows are padded to a multiple of 8 floats: the padding lanes of the last group are never read
        float4 *o = (float4 *)(tmp + ((size_t)blockIdx.z * nh + oyA) * tpitch + xb);
#pragma unroll
        for (int h = 0; h < NW; h++) o[h] = make_float4(tA[4 * h], tA[4 * h + 1], tA[4 * h + 2], tA[4 * h + 3]);
        if (hasB) {
            o = (float4 *)((float *)o + tpitch);
#pragma unroll
            for (int h = 0; h < NW; h++) o[h] = make_float4(tB[4 * h], tB[4 * h + 1], tB[4 * h + 2], tB[4 * h + 3]);
        }
    }
}

__device__ __forceinline__ uint8_t clamp_round_u8(float t) {
    // sample.rs: NumCast::from(FloatNearest(clamp(t, 0.0, 255.0))): clamp, then f32::round (half away from zero)
    t = t < 0.0f ? 0.0f : (t > 255.0f ? 255.0f : t);
    return (uint8_t)roundf(t);
}

__global__ void __launch_bounds__(128)
resize_horizontal_kernel(const float *__restrict__ tmp, uint32_t tpitch, const uint32_t *__restrict__ left,
                         const uint32_t *__restrict__ cnt, const float *__restrict__ wsT, uint32_t nw, uint32_t nh,
                         uint8_t *__restrict__ out, uint32_t n0) {
    const uint32_t ox = blockIdx.x * blockDim.x + threadIdx.x;
    if (ox >= nw) return;
    const uint32_t y = blockIdx.y, n = blockIdx.z + n0;
    const uint32_t c = cnt[ox];
    const float *__restrict__ row = tmp + ((size_t)blockIdx.z * nh + y) * tpitch + (size_t)left[ox] * 3;
    const float *__restrict__ w = wsT + ox;
    float t0 = 0.0f, t1 = 0.0f, t2 = 0.0f;
#pragma unroll 4
    for (uint32_t i = 0; i < c; i++) {
        const float wi = __ldg(w + (size_t)i * nw);
        t0 = __fadd_rn(t0, __fmul_rn(__ldg(row + 3 * i), wi));
        t1 = __fadd_rn(t1, __fmul_rn(__ldg(row + 3 * i + 1), wi));
        t2 = __fadd_rn(t2, __fmul_rn(__ldg(row + 3 * i + 2), wi));
    }
    uint8_t *o = out + (((size_t)n * nh + y) * nw + ox) * 3;
    o[0] = clamp_round_u8(t0);
    o[1] = clamp_round_u8(t1);
    o[2] = clamp_round_u8(t2);
}

// Row-major intermediate, windows of neighbouring outputs far apart: stage the row in shared memory first.
__device__ __forceinline__ uint32_t skew(uint32_t e) { return e + (e >> 5); }

// WS: the tap-major weights are staged too — with few blocks (the 64 rows of a single photo -> tile) every thread is a chain of
// dependent loads per tap, and a shared-memory load is ten times shorter than one from L2.
template <bool WS>
__global__ void __launch_bounds__(512)
resize_horizontal_smem_kernel(const float *__restrict__ tmp, uint32_t tpitch, uint32_t row_floats, const uint32_t *__restrict__ left,
                              const uint32_t *__restrict__ cnt, const float *__restrict__ wsT, uint32_t nw, uint32_t nh,
                              uint32_t w_words, uint32_t w_off, uint8_t *__restrict__ out, uint32_t n0) {
    extern __shared__ float srow[];
    const uint32_t y = blockIdx.x, n = blockIdx.y + n0;
    const float *__restrict__ row = tmp + ((size_t)blockIdx.y * nh + y) * tpitch;
    for (uint32_t e = threadIdx.x * 4; e < row_floats; e += blockDim.x * 4) {  // tpitch is a multiple of 8: whole float4s exist
        const float4 v = __ldg((const float4 *)(row + e));
        const uint32_t s = skew(e);  // e % 4 == 0: the four words share one skew
        srow[s] = v.x;
        srow[s + 1] = v.y;
        srow[s + 2] = v.z;
        srow[s + 3] = v.w;
    }
    float *sw = srow + w_off;  // [taps][nw]
    if (WS)
        for (uint32_t e = threadIdx.x; e < w_words; e += blockDim.x) sw[e] = __ldg(wsT + e);
    __syncthreads();
    for (uint32_t ox = threadIdx.x; ox < nw; ox += blockDim.x) {
        const uint32_t c = cnt[ox];
        const uint32_t base = left[ox] * 3;
        const float *__restrict__ w = (WS ? sw : wsT) + ox;
        float t0 = 0.0f, t1 = 0.0f, t2 = 0.0f;
#pragma unroll 4
        for (uint32_t i = 0; i < c; i++) {
            const float wi = WS ? w[(size_t)i * nw] : __ldg(w + (size_t)i * nw);
            const uint32_t e = base + 3 * i;
            t0 = __fadd_rn(t0, __fmul_rn(srow[skew(e)], wi));
            t1 = __fadd_rn(t1, __fmul_rn(srow[skew(e + 1)], wi));
            t2 = __fadd_rn(t2, __fmul_rn(srow[skew(e + 2)], wi));
        }
        uint8_t *o = out + (((size_t)n * nh + y) * nw + ox) * 3;
        o[0] = clamp_round_u8(t0);
        o[1] = clamp_round_u8(t1);
        o[2] = clamp_round_u8(t2);
    }
}

// Transposed intermediate [x*3+c][oy] (np floats per line): one block per (image, output column), lanes along oy, so a
// tap is three coalesced loads and the weight is uniform across the block.
__global__ void __launch_bounds__(128)
resize_horizontal_t_kernel(const float *__restrict__ tmpT, uint32_t tpitch, uint32_t np, const uint32_t *__restrict__ left,
                           const uint32_t *__restrict__ cnt, const float *__restrict__ ws, uint32_t wpitch, uint32_t nw, uint32_t nh,
                           uint8_t *__restrict__ out, uint32_t n0) {
    const uint32_t oy = blockIdx.y * blockDim.x + threadIdx.x;
    if (oy >= nh) return;
    const uint32_t ox = blockIdx.x, n = blockIdx.z + n0;
    const uint32_t c = cnt[ox];
    const float *__restrict__ col = tmpT + ((size_t)blockIdx.z * tpitch + (size_t)left[ox] * 3) * np + oy;
    const float *__restrict__ w = ws + (size_t)ox * wpitch;
    float t0 = 0.0f, t1 = 0.0f, t2 = 0.0f;
#pragma unroll 8
    for (uint32_t i = 0; i < c; i++, col += 3 * (size_t)np) {
        const float wi = __ldg(w + i);
        t0 = __fadd_rn(t0, __fmul_rn(__ldg(col), wi));
        t1 = __fadd_rn(t1, __fmul_rn(__ldg(col + np), wi));
        t2 = __fadd_rn(t2, __fmul_rn(__ldg(col + 2 * (size_t)np), wi));
    }
    uint8_t *o = out + (((size_t)n * nh + oy) * nw + ox) * 3;
    o[0] = clamp_round_u8(t0);
    o[1] = clamp_round_u8(t1);
    o[2] = clamp_round_u8(t2);
}

__global__ void __launch_bounds__(256)
resize_copy_kernel(const uint8_t *__restrict__ src, size_t img_bytes, uint32_t row_stride, size_t base_off, uint32_t row_bytes,
                   uint32_t ch, uint8_t *__restrict__ out) {
    const uint32_t xb = blockIdx.x * blockDim.x + threadIdx.x;
    if (xb >= row_bytes) return;
    const uint32_t y = blockIdx.y, n = blockIdx.z;
    out[((size_t)n * ch + y) * row_bytes + xb] = __ldg(src + (size_t)n * img_bytes + base_off + (size_t)y * row_stride + xb);
}

// ---- launcher -------------------------------------------------------------------------------------------------------------
struct emo_resize_state {
    emo_resize_axis v, h;
    uint32_t *d_tab = nullptr;  // device: left_v | cnt_v | left_h | cnt_h | ws_v [nh][pv] | -2^23 ws_v | ws_h tap-major [ph][nw] | ws_h [nw][ph]
    size_t d_tab_cap = 0;
    bool uploaded = false;
    uint32_t pair_span = 0;  // longest run of source rows two neighbouring output rows cover together; 0: windows not monotone
    float *tmp = nullptr;
    size_t tmp_cap = 0;
};

void emo_resize_state_free(emo_resize_state *s) {
    if (!s) return;
    cudaFree(s->d_tab);
    cudaFree(s->tmp);
    delete s;
}

int emo_launch_resize(emo_ctx *ctx, const uint8_t *images, uint32_t n, uint32_t img_w, uint32_t img_h, uint32_t x0, uint32_t y0,
                      uint32_t cw, uint32_t ch, uint32_t nw, uint32_t nh, uint8_t *out) {
    const uint32_t row_stride = img_w * 3, row_bytes = cw * 3;
    const size_t img_bytes = (size_t)img_w * img_h * 3;
    const size_t base_off = ((size_t)y0 * img_w + x0) * 3;
    if (nw == cw && nh == ch) {  // sample.rs resize(): same dimensions -> copy
        for (uint32_t z0 = 0; z0 < n; z0 += 32768) {
            const uint32_t nz = n - z0 < 32768 ? n - z0 : 32768;
            resize_copy_kernel<<<dim3((row_bytes + 255) / 256, ch, nz), 256, 0, ctx->stream>>>(
                images + (size_t)z0 * img_bytes, img_bytes, row_stride, base_off, row_bytes, ch, out + (size_t)z0 * ch * row_bytes);
            EMO_LAUNCH_CHECK(ctx);
        }
        return EMO_OK;
    }
    if (!ctx->resize) ctx->resize = new emo_resize_state();
    emo_resize_state &st = *ctx->resize;
    if (st.v.n_in != ch || st.v.n_out != nh) {
        resize_axis(ch, nh, st.v);
        st.uploaded = false;
        st.pair_span = 1;
        for (uint32_t o = 0; o < nh && st.pair_span; o += 2) {
            const uint32_t lA = st.v.left[o], eA = lA + st.v.cnt[o];
            const uint32_t lB = o + 1 < nh ? st.v.left[o + 1] : eA, eB = o + 1 < nh ? lB + st.v.cnt[o + 1] : eA;
            if (lA > lB || eA > eB) st.pair_span = 0;
            else if (eB - lA > st.pair_span) st.pair_span = eB - lA;
        }
    }
    if (st.h.n_in != cw || st.h.n_out != nw) { resize_axis(cw, nw, st.h); st.uploaded = false; }
    const uint32_t pv = (st.v.pitch + 3) / 4 * 4, ph = st.h.pitch;  // vertical weight rows padded for 128-bit loads
    // every section starts on a 16-byte boundary
    auto up4 = [](size_t x) { return (x + 3) / 4 * 4; };
    const size_t off_lv = 0, off_cv = up4(off_lv + nh), off_lh = up4(off_cv + nh), off_ch = up4(off_lh + nw), off_wv = up4(off_ch + nw),
                 off_nv = off_wv + (size_t)nh * pv, off_wh = off_nv + (size_t)nh * pv, off_whr = up4(off_wh + (size_t)ph * nw),
                 words = off_whr + (size_t)ph * nw;
    int rc;
    if (!st.uploaded) {
        if ((rc = emo_ensure(ctx, (void **)&st.d_tab, &st.d_tab_cap, words * 4))) return rc;
        std::vector<uint32_t> host(words, 0u);
        memcpy(host.data() + off_lv, st.v.left.data(), (size_t)nh * 4);
        memcpy(host.data() + off_cv, st.v.cnt.data(), (size_t)nh * 4);
        memcpy(host.data() + off_lh, st.h.left.data(), (size_t)nw * 4);
        memcpy(host.data() + off_ch, st.h.cnt.data(), (size_t)nw * 4);
        float *wv = (float *)(host.data() + off_wv), *nv = (float *)(host.data() + off_nv);
        for (uint32_t o = 0; o < nh; o++)
            for (uint32_t i = 0; i < st.v.pitch; i++) {
                const float wi = st.v.ws[(size_t)o * st.v.pitch + i];
                wv[(size_t)o * pv + i] = wi;
                nv[(size_t)o * pv + i] = wi * -8388608.0f;  // exact (a power of two), the addend of the kernel's FFMA
            }
        float *wt = (float *)(host.data() + off_wh);  // tap-major copy for the row-major horizontal kernel
        for (uint32_t o = 0; o < nw; o++)
            for (uint32_t i = 0; i < ph; i++) wt[(size_t)i * nw + o] = st.h.ws[(size_t)o * ph + i];
        memcpy(host.data() + off_whr, st.h.ws.data(), (size_t)nw * ph * 4);
        // pageable source: the runtime stages it before returning, `host` may go out of scope
        EMO_CK(cudaMemcpyAsync(st.d_tab, host.data(), words * 4, cudaMemcpyHostToDevice, ctx->stream));
        EMO_CK(cudaStreamSynchronize(ctx->stream));
        st.uploaded = true;
    }
    const uint32_t *d_lv = st.d_tab + off_lv, *d_cv = st.d_tab + off_cv, *d_lh = st.d_tab + off_lh, *d_ch = st.d_tab + off_ch;
    const float *d_wv = (const float *)(st.d_tab + off_wv), *d_nv = (const float *)(st.d_tab + off_nv),
                *d_wh = (const float *)(st.d_tab + off_wh), *d_whr = (const float *)(st.d_tab + off_whr);
    const uint32_t tpitch = (row_bytes + 7) / 8 * 8;
    // Transposed intermediate (see the header comment) for batches of photo -> tile reductions: both axes shrink 8x or more
    // (48+ taps amortise the four scattered stores of the vertical pass), at least a warp of output rows, and enough outputs
    // that the horizontal pass is throughput-bound.  Measured (tools/bench_resize.py, ncu launch lists under profiles/):
    // 64 x 2048^2 -> 64^2 horizontal pass 615 -> 90 us; but 8192^2 -> 4096^2 (13 taps) 2.7x slower and a single
    // 4000x3000 photo (4096 outputs, latency-bound, the row-major layout keeps a thread's taps in one L1 line) 1.3x slower;
    // such geometries take the shared-memory staged horizontal pass below instead.
    const bool transposed = cw >= 8 * (uint64_t)nw && ch >= 8 * (uint64_t)nh && nh >= 32 && (uint64_t)n * nh * nw >= 65536;
    const uint32_t np = (nh + 31) / 32 * 32;
    // images per pass: bounded by the f32 intermediate (<= 1 GiB) and the grid's z extent
    const size_t tmp_per_img = (size_t)(transposed ? np : nh) * tpitch * 4;
    uint32_t per_pass = (uint32_t)((1ull << 30) / tmp_per_img);
    if (per_pass < 1) per_pass = 1;
    if (per_pass > n) per_pass = n;
    if (per_pass > 32768) per_pass = 32768;
    if ((rc = emo_ensure(ctx, (void **)&st.tmp, &st.tmp_cap, tmp_per_img * per_pass))) return rc;
    const bool aligned = ((uintptr_t)images % 8 == 0) && (img_bytes % 8 == 0) && (row_stride % 8 == 0) && (base_off % 8 == 0);
    const uint32_t groups = tpitch / 8;
    // row-major intermediate with outputs 4+ source pixels apart: stage each row in shared memory (skewed), if it fits
    const size_t smem_row = ((size_t)tpitch + tpitch / 32 + 8) * 4;
    const bool staged = !transposed && cw >= 4 * (uint64_t)nw && smem_row <= 200 * 1024;
    // few blocks (at most one per SM): the weights go to shared memory as well, if they fit
    const size_t w_words = (size_t)ph * nw, w_off = (smem_row / 4 + 3) / 4 * 4;
    const bool staged_w = staged && (uint64_t)nh * n <= (uint64_t)ctx->sm_count && (w_off + w_words) * 4 <= 200 * 1024;
    const size_t smem_h = staged_w ? (w_off + w_words) * 4 : smem_row;
    if (staged && smem_h > 48 * 1024)
        EMO_CK(cudaFuncSetAttribute(staged_w ? resize_horizontal_smem_kernel<true> : resize_horizontal_smem_kernel<false>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (uint32_t z0 = 0; z0 < n; z0 += per_pass) {
        const uint32_t nz = n - z0 < per_pass ? n - z0 : per_pass;
        const dim3 gv((groups + 255) / 256, nh, nz);
        const size_t pair_smem = (size_t)(st.pair_span + 3) / 4 * 4 * sizeof(float4);
        static const bool pair_off = getenv("EMO_RESIZE_PAIR") && atoi(getenv("EMO_RESIZE_PAIR")) == 0;  // A-B switch
        if (aligned && st.pair_span && pair_smem <= 48 * 1024 && !pair_off) {
            // 8 bytes per thread and loads one group ahead keep a full grid busy; a grid of at most one wave (a single photo: 192
            // blocks) takes 4 bytes per thread — twice the warps — and three groups in flight
            static const int depth_env = getenv("EMO_RESIZE_DEPTH") ? atoi(getenv("EMO_RESIZE_DEPTH")) : 0;  // tuning overrides
            static const int nb_env = getenv("EMO_RESIZE_NB") ? atoi(getenv("EMO_RESIZE_NB")) : 0;
            const bool small = (uint64_t)((groups + 255) / 256) * ((nh + 1) / 2) * nz <= 4ull * ctx->sm_count;
            const int depth = depth_env ? depth_env : (small ? 3 : 1), nb = nb_env ? nb_env : (small ? 4 : 8);
            const uint32_t cols = tpitch / nb;  // threads per row
            const uint32_t bs = small ? 128 : (cols >= 256 ? 256 : (cols + 31) / 32 * 32);
            const dim3 g2((cols + bs - 1) / bs, (nh + 1) / 2, nz);
#define EMO_VERTICAL2(T, P, B)                                                                                                     \
    resize_vertical2_kernel<T, P, B><<<g2, bs, pair_smem, ctx->stream>>>(images, img_bytes, row_stride, base_off, row_bytes, d_lv, d_cv, \
                                                                         d_wv, pv, st.tmp, tpitch, nh, np, z0)
            if (transposed) {
                if (nb == 4) { if (depth == 1) EMO_VERTICAL2(true, 1, 4); else EMO_VERTICAL2(true, 3, 4); }
                else if (depth == 1) EMO_VERTICAL2(true, 1, 8); else EMO_VERTICAL2(true, 3, 8);
            } else {
                if (nb == 4) { if (depth == 1) EMO_VERTICAL2(false, 1, 4); else EMO_VERTICAL2(false, 3, 4); }
                else if (depth == 1) EMO_VERTICAL2(false, 1, 8); else EMO_VERTICAL2(false, 3, 8);
            }
#undef EMO_VERTICAL2
        } else
#define EMO_VERTICAL(A, T)                                                                                                         \
    resize_vertical_kernel<A, T><<<gv, 256, 0, ctx->stream>>>(images, img_bytes, row_stride, base_off, row_bytes, d_lv, d_cv, d_wv, d_nv, pv, \
                                                             st.tmp, tpitch, nh, np, z0)
        if (transposed) {
            if (aligned) EMO_VERTICAL(true, true); else EMO_VERTICAL(false, true);
        } else {
            if (aligned) EMO_VERTICAL(true, false); else EMO_VERTICAL(false, false);
        }
#undef EMO_VERTICAL
        EMO_LAUNCH_CHECK(ctx);
        if (transposed) {
            const uint32_t bt = nh >= 128 ? 128 : (nh + 31) / 32 * 32;
            resize_horizontal_t_kernel<<<dim3(nw, (nh + bt - 1) / bt, nz), bt, 0, ctx->stream>>>(st.tmp, tpitch, np, d_lh, d_ch, d_whr, ph,
                                                                                               nw, nh, out, z0);
        } else if (staged) {
            // the staging loop wants many loads in flight whatever the number of outputs: a thread per 32 floats of the row
            uint32_t bt = (row_bytes / 32 + 31) / 32 * 32;
            if (bt < (nw + 31) / 32 * 32) bt = (nw + 31) / 32 * 32;
            bt = bt > 512 ? 512 : (bt < 32 ? 32 : bt);
            if (staged_w)
                resize_horizontal_smem_kernel<true><<<dim3(nh, nz), bt, smem_h, ctx->stream>>>(st.tmp, tpitch, row_bytes, d_lh, d_ch, d_wh, nw, nh,
                                                                                              (uint32_t)w_words, (uint32_t)w_off, out, z0);
            else
                resize_horizontal_smem_kernel<false><<<dim3(nh, nz), bt, smem_h, ctx->stream>>>(st.tmp, tpitch, row_bytes, d_lh, d_ch, d_wh, nw,
                                                                                               nh, 0u, 0u, out, z0);
        } else {
            resize_horizontal_kernel<<<dim3((nw + 127) / 128, nh, nz), 128, 0, ctx->stream>>>(st.tmp, tpitch, d_lh, d_ch, d_wh, nw, nh,
                                                                                            out, z0);
        }
        EMO_LAUNCH_CHECK(ctx);
    }
    return EMO_OK;
}
