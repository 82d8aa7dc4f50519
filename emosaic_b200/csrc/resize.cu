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
// that order; the parallelism is across output elements.  The multiply and the add are issued with __fmul_rn /
// __fadd_rn so ptxas cannot fuse them.
//
// The tap tables (left, count, normalised Lanczos3 weights: O(nw + nh) numbers, each needing the platform's libm sinf
// exactly as f32::sin does in the reference) are computed on the host in resize_axis() and cached per geometry; all
// O(pixels) arithmetic runs here.
//
//   resize_vertical_kernel<ALIGNED>: one thread per (image, output row, 4 consecutive bytes of the view row).  Rows of
//     the view are read as 32-bit words when the geometry keeps them 4-byte aligned, bytes otherwise; a byte becomes an
//     f32 exactly with PRMT (0x4B0000bb = 2^23 + b) and one FADD.  Writes one float4 per thread into tmp (row pitch
//     padded to 4 floats).  The ~6x re-read of a source row by neighbouring output rows is served by L1/L2.
//   resize_horizontal_kernel: one thread per (image, row, output pixel), three channel accumulators, weights stored
//     tap-major ([tap][ox]) so that lanes read consecutive words.
//   resize_copy_kernel: the "(nwidth, nheight) == image.dimensions()" early return of resize() — a plain copy.
#include <math.h>

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

// ---- device ---------------------------------------------------------------------------------------------------------------
// exact u8 -> f32: PRMT builds 0x4B0000bb = 2^23 + b, one FADD removes the 2^23
template <int K>
__device__ __forceinline__ float byte_f32(uint32_t v) {
    return __fadd_rn(__uint_as_float(__byte_perm(v, 0x4B000000u, 0x7540 | K)), -8388608.0f);
}

template <bool ALIGNED>
__global__ void __launch_bounds__(256)
resize_vertical_kernel(const uint8_t *__restrict__ src, size_t img_bytes, uint32_t row_stride, size_t base_off, uint32_t row_bytes,
                       const uint32_t *__restrict__ left, const uint32_t *__restrict__ cnt, const float *__restrict__ ws,
                       uint32_t wpitch, float *__restrict__ tmp, uint32_t tpitch, uint32_t nh, uint32_t n0) {
    const uint32_t xb = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (xb >= row_bytes) return;
    const uint32_t oy = blockIdx.y, n = blockIdx.z + n0;
    const uint32_t c = cnt[oy];
    const float *__restrict__ w = ws + (size_t)oy * wpitch;
    const uint8_t *p = src + (size_t)n * img_bytes + base_off + (size_t)left[oy] * row_stride + xb;
    const uint32_t valid = row_bytes - xb;  // bytes of this thread's group that belong to the view (>= 1)
    float t0 = 0.0f, t1 = 0.0f, t2 = 0.0f, t3 = 0.0f;
    if (ALIGNED && valid >= 4) {
#pragma unroll 4
        for (uint32_t i = 0; i < c; i++, p += row_stride) {
            const uint32_t v = __ldg((const uint32_t *)p);
            const float wi = __ldg(w + i);
            t0 = __fadd_rn(t0, __fmul_rn(byte_f32<0>(v), wi));
            t1 = __fadd_rn(t1, __fmul_rn(byte_f32<1>(v), wi));
            t2 = __fadd_rn(t2, __fmul_rn(byte_f32<2>(v), wi));
            t3 = __fadd_rn(t3, __fmul_rn(byte_f32<3>(v), wi));
        }
    } else {
#pragma unroll 2
        for (uint32_t i = 0; i < c; i++, p += row_stride) {
            uint32_t v = __ldg(p);
            if (valid > 1) v |= (uint32_t)__ldg(p + 1) << 8;
            if (valid > 2) v |= (uint32_t)__ldg(p + 2) << 16;
            if (valid > 3) v |= (uint32_t)__ldg(p + 3) << 24;
            const float wi = __ldg(w + i);
            t0 = __fadd_rn(t0, __fmul_rn(byte_f32<0>(v), wi));
            t1 = __fadd_rn(t1, __fmul_rn(byte_f32<1>(v), wi));
            t2 = __fadd_rn(t2, __fmul_rn(byte_f32<2>(v), wi));
            t3 = __fadd_rn(t3, __fmul_rn(byte_f32<3>(v), wi));
        }
    }
    // tmp rows are padded to a multiple of 4 floats: the padding lanes of the last group hold sums of zero bytes
    *(float4 *)(tmp + ((size_t)blockIdx.z * nh + oy) * tpitch + xb) = make_float4(t0, t1, t2, t3);
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
    uint32_t *d_tab = nullptr;  // device: left_v | cnt_v | left_h | cnt_h | ws_v [nh][pv] | ws_h transposed [ph][nw]
    size_t d_tab_cap = 0;
    bool uploaded = false;
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
    if (st.v.n_in != ch || st.v.n_out != nh) { resize_axis(ch, nh, st.v); st.uploaded = false; }
    if (st.h.n_in != cw || st.h.n_out != nw) { resize_axis(cw, nw, st.h); st.uploaded = false; }
    const uint32_t pv = st.v.pitch, ph = st.h.pitch;
    const size_t off_lv = 0, off_cv = off_lv + nh, off_lh = off_cv + nh, off_ch = off_lh + nw, off_wv = off_ch + nw,
                 off_wh = off_wv + (size_t)nh * pv, words = off_wh + (size_t)ph * nw;
    int rc;
    if (!st.uploaded) {
        if ((rc = emo_ensure(ctx, (void **)&st.d_tab, &st.d_tab_cap, words * 4))) return rc;
        std::vector<uint32_t> host(words);
        memcpy(host.data() + off_lv, st.v.left.data(), (size_t)nh * 4);
        memcpy(host.data() + off_cv, st.v.cnt.data(), (size_t)nh * 4);
        memcpy(host.data() + off_lh, st.h.left.data(), (size_t)nw * 4);
        memcpy(host.data() + off_ch, st.h.cnt.data(), (size_t)nw * 4);
        memcpy(host.data() + off_wv, st.v.ws.data(), (size_t)nh * pv * 4);
        float *wt = (float *)(host.data() + off_wh);
        for (uint32_t o = 0; o < nw; o++)
            for (uint32_t i = 0; i < ph; i++) wt[(size_t)i * nw + o] = st.h.ws[(size_t)o * ph + i];
        // pageable source: the runtime stages it before returning, `host` may go out of scope
        EMO_CK(cudaMemcpyAsync(st.d_tab, host.data(), words * 4, cudaMemcpyHostToDevice, ctx->stream));
        EMO_CK(cudaStreamSynchronize(ctx->stream));
        st.uploaded = true;
    }
    const uint32_t *d_lv = st.d_tab + off_lv, *d_cv = st.d_tab + off_cv, *d_lh = st.d_tab + off_lh, *d_ch = st.d_tab + off_ch;
    const float *d_wv = (const float *)(st.d_tab + off_wv), *d_wh = (const float *)(st.d_tab + off_wh);
    const uint32_t tpitch = (row_bytes + 3) / 4 * 4;
    // images per pass: bounded by the f32 intermediate (<= 1 GiB) and the grid's z extent
    const size_t tmp_per_img = (size_t)nh * tpitch * 4;
    uint32_t per_pass = (uint32_t)((1ull << 30) / tmp_per_img);
    if (per_pass < 1) per_pass = 1;
    if (per_pass > n) per_pass = n;
    if (per_pass > 32768) per_pass = 32768;
    if ((rc = emo_ensure(ctx, (void **)&st.tmp, &st.tmp_cap, tmp_per_img * per_pass))) return rc;
    const bool aligned = ((uintptr_t)images % 4 == 0) && (img_bytes % 4 == 0) && (row_stride % 4 == 0) && (base_off % 4 == 0);
    const uint32_t groups = tpitch / 4;
    for (uint32_t z0 = 0; z0 < n; z0 += per_pass) {
        const uint32_t nz = n - z0 < per_pass ? n - z0 : per_pass;
        const dim3 gv((groups + 255) / 256, nh, nz);
        if (aligned)
            resize_vertical_kernel<true><<<gv, 256, 0, ctx->stream>>>(images, img_bytes, row_stride, base_off, row_bytes, d_lv, d_cv,
                                                                    d_wv, pv, st.tmp, tpitch, nh, z0);
        else
            resize_vertical_kernel<false><<<gv, 256, 0, ctx->stream>>>(images, img_bytes, row_stride, base_off, row_bytes, d_lv, d_cv,
                                                                     d_wv, pv, st.tmp, tpitch, nh, z0);
        EMO_LAUNCH_CHECK(ctx);
        resize_horizontal_kernel<<<dim3((nw + 127) / 128, nh, nz), 128, 0, ctx->stream>>>(st.tmp, tpitch, d_lh, d_ch, d_wh, nw, nh,
                                                                                        out, z0);
        EMO_LAUNCH_CHECK(ctx);
    }
    return EMO_OK;
}
