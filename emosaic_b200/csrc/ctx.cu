// ctx.cu — context management and the host-pointer halves of the C ABI (include/emosaic_cuda.h).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

static thread_local std::string g_last_error;

void emo_set_error(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

extern "C" {

int emo_abi_version(void) { return EMO_ABI_VERSION; }
const char *emo_last_error(void) { return g_last_error.c_str(); }

int emo_create(int device, emo_ctx **out) {
    EMO_REQUIRE(out != nullptr, EMO_ERR_ARG, "emo_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        emo_set_error("emo_create: no CUDA device (%s); this library has no CPU path",
                      e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return EMO_ERR_NO_DEVICE;
    }
    EMO_REQUIRE(device >= 0 && device < n, EMO_ERR_ARG, "emo_create: device %d out of range [0,%d)", device, n);
    EMO_CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    EMO_CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        emo_set_error("emo_create: device %d (%s, cc %d.%d) is not sm_100; kernels are built for sm_100a only",
                      device, prop.name, prop.major, prop.minor);
        return EMO_ERR_NO_DEVICE;
    }
    emo_ctx *ctx = new emo_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->cc_major = prop.major;
    ctx->cc_minor = prop.minor;
    const int rc = [&]() -> int {
        EMO_CK(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
        EMO_CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        ctx->stream = ctx->own_stream;
        EMO_CK(cudaEventCreate(&ctx->ev_start));
        EMO_CK(cudaEventCreate(&ctx->ev_stop));
        for (int i = 0; i < 4; i++) EMO_CK(cudaEventCreateWithFlags(&ctx->ev_pipe[i], cudaEventDisableTiming));
        EMO_CK(cudaMalloc(&ctx->err_flag, sizeof(int)));
        EMO_CK(cudaMemset(ctx->err_flag, 0, sizeof(int)));
        return EMO_OK;
    }();
    if (rc != EMO_OK) {  // a half-built ctx is released here; the error message of the failing call stays
        const std::string msg = g_last_error;
        emo_destroy(ctx);
        g_last_error = msg;
        return rc;
    }
    *out = ctx;
    return EMO_OK;
}

void emo_destroy(emo_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    emo_comm_release(ctx);
    cudaFree(ctx->comm_hdr);
    cudaFree(ctx->cand);
    cudaFree(ctx->lib_px);
    cudaFree(ctx->lut);
    cudaFree(ctx->lut16);
    cudaFree(ctx->idx_seeded);
    cudaFree(ctx->idx_slot_of_tile);
    cudaFree(ctx->idx_entry);
    if (ctx->idx_count_host) cudaFreeHost((void *)ctx->idx_count_host);
    if (ctx->idx_count_ev) cudaEventDestroy(ctx->idx_count_ev);
    cudaFree(ctx->keys);
    cudaFree(ctx->qvec);
    cudaFree(ctx->err_flag);
    for (int i = 0; i < 6; i++) cudaFree(ctx->stage[i]);
    cudaFree(ctx->tint.lut);
    cudaFree(ctx->tint.excv);
    cudaFree(ctx->tint.excm);
    cudaFree(ctx->tint.cadd);
    cudaFree(ctx->tint.meta);
    emo_resize_state_free(ctx->resize);
    for (int i = 0; i < 4; i++)
        if (ctx->ev_pipe[i]) cudaEventDestroy(ctx->ev_pipe[i]);
    for (cudaEvent_t e : ctx->marks)
        if (e) cudaEventDestroy(e);
    if (ctx->ev_start) cudaEventDestroy(ctx->ev_start);
    if (ctx->ev_stop) cudaEventDestroy(ctx->ev_stop);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    delete ctx;
}

int emo_set_stream(emo_ctx *ctx, void *cuda_stream) {
    EMO_REQUIRE(ctx, EMO_ERR_ARG, "emo_set_stream: ctx is NULL");
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return EMO_OK;
}

int emo_sync(emo_ctx *ctx) {
    EMO_REQUIRE(ctx, EMO_ERR_ARG, "emo_sync: ctx is NULL");
    EMO_CK(cudaSetDevice(ctx->device));
    EMO_CK(cudaStreamSynchronize(ctx->stream));
    EMO_CK(cudaStreamSynchronize(ctx->copy_stream));
    return emo_check_device_flag(ctx);
}

int emo_device_info(emo_ctx *ctx, char *name, size_t name_cap, int *sm_count, int *cc_major, int *cc_minor) {
    EMO_REQUIRE(ctx, EMO_ERR_ARG, "emo_device_info: ctx is NULL");
    cudaDeviceProp prop;
    EMO_CK(cudaGetDeviceProperties(&prop, ctx->device));
    if (name && name_cap) {
        strncpy(name, prop.name, name_cap - 1);
        name[name_cap - 1] = 0;
    }
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return EMO_OK;
}

uint64_t emo_launch_count(emo_ctx *ctx) { return ctx ? ctx->launches : 0; }

int emo_timer_start(emo_ctx *ctx) {
    EMO_REQUIRE(ctx, EMO_ERR_ARG, "emo_timer_start: ctx is NULL");
    EMO_CK(cudaEventRecord(ctx->ev_start, ctx->stream));
    return EMO_OK;
}

int emo_timer_stop(emo_ctx *ctx, float *elapsed_ms) {
    EMO_REQUIRE(ctx && elapsed_ms, EMO_ERR_ARG, "emo_timer_stop: NULL argument");
    EMO_CK(cudaEventRecord(ctx->ev_stop, ctx->stream));
    EMO_CK(cudaEventSynchronize(ctx->ev_stop));
    EMO_CK(cudaEventElapsedTime(elapsed_ms, ctx->ev_start, ctx->ev_stop));
    return EMO_OK;
}

int emo_mark(emo_ctx *ctx, uint32_t slot) {
    EMO_REQUIRE(ctx && slot < 65536, EMO_ERR_ARG, "emo_mark: bad argument");
    if (ctx->marks.size() <= slot) ctx->marks.resize(slot + 1, nullptr);
    if (!ctx->marks[slot]) EMO_CK(cudaEventCreate(&ctx->marks[slot]));
    EMO_CK(cudaEventRecord(ctx->marks[slot], ctx->stream));
    return EMO_OK;
}

int emo_mark_elapsed(emo_ctx *ctx, uint32_t a, uint32_t b, float *elapsed_ms) {
    EMO_REQUIRE(ctx && elapsed_ms && a < ctx->marks.size() && b < ctx->marks.size() && ctx->marks[a] && ctx->marks[b],
                EMO_ERR_ARG, "emo_mark_elapsed: unknown mark");
    EMO_CK(cudaEventSynchronize(ctx->marks[b]));
    EMO_CK(cudaEventElapsedTime(elapsed_ms, ctx->marks[a], ctx->marks[b]));
    return EMO_OK;
}

int emo_dev_alloc(emo_ctx *ctx, size_t bytes, void **out) {
    EMO_REQUIRE(ctx && out, EMO_ERR_ARG, "emo_dev_alloc: NULL argument");
    EMO_CK(cudaSetDevice(ctx->device));
    EMO_CK(cudaMalloc(out, bytes ? bytes : 1));
    return EMO_OK;
}
int emo_dev_free(emo_ctx *ctx, void *p) {
    EMO_REQUIRE(ctx, EMO_ERR_ARG, "emo_dev_free: ctx is NULL");
    EMO_CK(cudaFree(p));
    return EMO_OK;
}
int emo_host_alloc(emo_ctx *ctx, size_t bytes, void **out) {
    EMO_REQUIRE(ctx && out, EMO_ERR_ARG, "emo_host_alloc: NULL argument");
    EMO_CK(cudaSetDevice(ctx->device));
    EMO_CK(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return EMO_OK;
}
int emo_host_free(emo_ctx *ctx, void *p) {
    EMO_REQUIRE(ctx, EMO_ERR_ARG, "emo_host_free: ctx is NULL");
    EMO_CK(cudaFreeHost(p));
    return EMO_OK;
}
int emo_copy_h2d(emo_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes) {
    EMO_REQUIRE(ctx && (bytes == 0 || (dst_dev && src_host)), EMO_ERR_ARG, "emo_copy_h2d: NULL argument");
    EMO_CK(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return EMO_OK;
}
int emo_copy_d2h(emo_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes) {
    EMO_REQUIRE(ctx && (bytes == 0 || (dst_host && src_dev)), EMO_ERR_ARG, "emo_copy_d2h: NULL argument");
    EMO_CK(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return EMO_OK;
}

}  // extern "C"

int emo_ensure(emo_ctx *ctx, void **p, size_t *cap, size_t bytes) {
    if (*cap >= bytes && *p) return EMO_OK;
    EMO_CK(cudaSetDevice(ctx->device));
    if (*p) {  // both streams may still be working on the old buffer (a pipelined host call that returned on an error)
        EMO_CK(cudaStreamSynchronize(ctx->stream));
        EMO_CK(cudaStreamSynchronize(ctx->copy_stream));
        EMO_CK(cudaFree(*p));
        *p = nullptr;
        *cap = 0;
    }
    size_t want = bytes < 256 ? 256 : bytes;
    cudaError_t e = cudaMalloc(p, want);
    if (e != cudaSuccess) {
        emo_set_error("device allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
        *p = nullptr;
        return EMO_ERR_OOM;
    }
    *cap = want;
    return EMO_OK;
}

int emo_check_device_flag(emo_ctx *ctx) {
    int flag = 0;
    EMO_CK(cudaMemcpy(&flag, ctx->err_flag, sizeof(int), cudaMemcpyDeviceToHost));
    if (flag) {
        EMO_CK(cudaMemset(ctx->err_flag, 0, sizeof(int)));
        emo_set_error("compose: item map holds 0 or an id beyond the library (reference: tileset.rs:131-143 get_tile -> None, "
                      "rendering.rs:198-206 panics)");
        return EMO_ERR_ARG;
    }
    return EMO_OK;
}

// ---------------------------------------------------------------------------------------
// C ABI: device-pointer variants (thin argument checks + launchers) and host-pointer variants
// ---------------------------------------------------------------------------------------
int emo_check_analyse_args(const void *tiles, uint64_t T, uint32_t ts, uint32_t dim, const void *out) {
    EMO_REQUIRE(T == 0 || (tiles && out), EMO_ERR_ARG, "analyse: NULL buffer");
    EMO_REQUIRE(ts >= 1 && ts <= 4096, EMO_ERR_ARG, "analyse: tile size %u outside [1,4096]", ts);
    // color.rs:18 "Rectangle dimensions must be positive": floor(ts/dim) == 0 panics in the reference
    EMO_REQUIRE(dim >= 1 && dim <= ts, EMO_ERR_ARG,
                "analyse: dim %u gives an empty cell for tile size %u (Rectangle dimensions must be positive)", dim, ts);
    return EMO_OK;
}

extern "C" {

int emo_analyse_dev(emo_ctx *ctx, const uint8_t *tiles, uint64_t T, uint32_t ts, uint32_t dim, uint8_t *out) {
    EMO_REQUIRE(ctx, EMO_ERR_ARG, "emo_analyse_dev: ctx is NULL");
    int rc = emo_check_analyse_args(tiles, T, ts, dim, out);
    if (rc) return rc;
    if (T == 0) return EMO_OK;
    EMO_CK(cudaSetDevice(ctx->device));
    return emo_launch_analyse(ctx, tiles, T, ts, dim, out);
}

int emo_analyse_fused_dev(emo_ctx *ctx, const uint8_t *tiles, uint64_t T, uint32_t ts, uint8_t *out1, uint8_t *out4) {
    EMO_REQUIRE(ctx, EMO_ERR_ARG, "emo_analyse_fused_dev: ctx is NULL");
    int rc = emo_check_analyse_args(tiles, T, ts, 2, out1);
    if (rc) return rc;
    EMO_REQUIRE(T == 0 || out4, EMO_ERR_ARG, "analyse_fused: out4 is NULL");
    EMO_REQUIRE(ts % 2 == 0, EMO_ERR_ARG, "Invalid tile size: Tile size must be divisible by 2");  // main.rs:612-615
    if (T == 0) return EMO_OK;
    EMO_CK(cudaSetDevice(ctx->device));
    return emo_launch_analyse_fused(ctx, tiles, T, ts, out1, out4);
}

}  // extern "C"

// Host-pointer analysis streams the library through two device slabs (~256 MB each): the H2D of slab k+1 runs
// on the copy stream while slab k is reduced, so a library larger than HBM (or than the caller wants to stage) works
// and the kernels hide behind PCIe.  dim2 == 0: single analysis into out1; otherwise the fused 1to1 + 4to1 pass.
int emo_analyse_host_impl(emo_ctx *ctx, const uint8_t *tiles, uint64_t T, uint32_t ts, uint32_t dim, uint8_t *out1, uint8_t *out4,
                          bool fused) {
    const size_t tile_b = (size_t)ts * ts * 3;
    const size_t out_per_tile = fused ? 15 : (size_t)dim * dim * 3;
    uint64_t slab_tiles = (256ull << 20) / tile_b;
    if (slab_tiles < 1) slab_tiles = 1;
    if (slab_tiles > T) slab_tiles = T;
    int rc;
    if ((rc = emo_ensure(ctx, &ctx->stage[0], &ctx->stage_cap[0], 2 * slab_tiles * tile_b))) return rc;
    if ((rc = emo_ensure(ctx, &ctx->stage[1], &ctx->stage_cap[1], 2 * slab_tiles * out_per_tile))) return rc;
    uint8_t *din[2] = {(uint8_t *)ctx->stage[0], (uint8_t *)ctx->stage[0] + slab_tiles * tile_b};
    uint8_t *dout[2] = {(uint8_t *)ctx->stage[1], (uint8_t *)ctx->stage[1] + slab_tiles * out_per_tile};
    uint32_t k = 0;
    for (uint64_t t0 = 0; t0 < T; t0 += slab_tiles, k++) {
        const uint64_t n = T - t0 < slab_tiles ? T - t0 : slab_tiles;
        const int b = k & 1;
        // slab buffer b is free once the kernel + D2H of slab k-2 (issued on the compute stream) are done
        if (k >= 2) EMO_CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_pipe[2 + b], 0));
        EMO_CK(cudaMemcpyAsync(din[b], tiles + t0 * tile_b, n * tile_b, cudaMemcpyHostToDevice, ctx->copy_stream));
        EMO_CK(cudaEventRecord(ctx->ev_pipe[b], ctx->copy_stream));
        EMO_CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_pipe[b], 0));
        if (fused) {
            uint8_t *d1 = dout[b], *d4 = dout[b] + n * 3;
            if ((rc = emo_launch_analyse_fused(ctx, din[b], n, ts, d1, d4))) return rc;
            EMO_CK(cudaMemcpyAsync(out1 + t0 * 3, d1, n * 3, cudaMemcpyDeviceToHost, ctx->stream));
            EMO_CK(cudaMemcpyAsync(out4 + t0 * 12, d4, n * 12, cudaMemcpyDeviceToHost, ctx->stream));
        } else {
            if ((rc = emo_launch_analyse(ctx, din[b], n, ts, dim, dout[b]))) return rc;
            EMO_CK(cudaMemcpyAsync(out1 + t0 * out_per_tile, dout[b], n * out_per_tile, cudaMemcpyDeviceToHost, ctx->stream));
        }
        EMO_CK(cudaEventRecord(ctx->ev_pipe[2 + b], ctx->stream));
    }
    EMO_CK(cudaStreamSynchronize(ctx->stream));
    EMO_CK(cudaStreamSynchronize(ctx->copy_stream));
    return EMO_OK;
}

extern "C" {

int emo_analyse(emo_ctx *ctx, const uint8_t *tiles, uint64_t T, uint32_t ts, uint32_t dim, uint8_t *out) {
    EMO_REQUIRE(ctx, EMO_ERR_ARG, "emo_analyse: ctx is NULL");
    int rc = emo_check_analyse_args(tiles, T, ts, dim, out);
    if (rc) return rc;
    if (T == 0) return EMO_OK;
    EMO_CK(cudaSetDevice(ctx->device));
    return emo_analyse_host_impl(ctx, tiles, T, ts, dim, out, nullptr, false);
}

int emo_analyse_fused(emo_ctx *ctx, const uint8_t *tiles, uint64_t T, uint32_t ts, uint8_t *out1, uint8_t *out4) {
    EMO_REQUIRE(ctx, EMO_ERR_ARG, "emo_analyse_fused: ctx is NULL");
    int rc = emo_check_analyse_args(tiles, T, ts, 2, out1);
    if (rc) return rc;
    EMO_REQUIRE(T == 0 || out4, EMO_ERR_ARG, "analyse_fused: out4 is NULL");
    EMO_REQUIRE(ts % 2 == 0, EMO_ERR_ARG, "Invalid tile size: Tile size must be divisible by 2");
    if (T == 0) return EMO_OK;
    EMO_CK(cudaSetDevice(ctx->device));
    return emo_analyse_host_impl(ctx, tiles, T, ts, 2, out1, out4, true);
}

// ---- Lanczos3 resize (image 0.25.2 imageops::resize; main.rs:595, tiles/utils.rs:188-189) -----------------------------
static int check_resize_args(emo_ctx *ctx, const void *images, uint32_t n, uint32_t img_w, uint32_t img_h, uint32_t x0, uint32_t y0,
                             uint32_t cw, uint32_t ch, uint32_t nw, uint32_t nh, const void *out) {
    EMO_REQUIRE(ctx, EMO_ERR_ARG, "resize: ctx is NULL");
    EMO_REQUIRE(n == 0 || (images && out), EMO_ERR_ARG, "resize: NULL buffer");
    EMO_REQUIRE(img_w >= 1 && img_h >= 1 && cw >= 1 && ch >= 1 && nw >= 1 && nh >= 1, EMO_ERR_ARG,
                "resize: empty image, view or output (%ux%u, view %ux%u -> %ux%u)", img_w, img_h, cw, ch, nw, nh);
    EMO_REQUIRE((uint64_t)x0 + cw <= img_w && (uint64_t)y0 + ch <= img_h, EMO_ERR_ARG,
                "resize: view (%u,%u,%u,%u) extends beyond the %ux%u image", x0, y0, cw, ch, img_w, img_h);
    if (img_w > (1u << 24) || img_h > 65535 || nh > 65535 || nw > (1u << 24)) {
        emo_set_error("resize: dimensions beyond 2^24 columns / 65535 rows are not supported (%ux%u -> %ux%u)", img_w, img_h, nw, nh);
        return EMO_ERR_UNSUPPORTED;
    }
    return EMO_OK;
}

int emo_resize_dev(emo_ctx *ctx, const uint8_t *images, uint32_t n, uint32_t img_w, uint32_t img_h, uint32_t x0, uint32_t y0,
                   uint32_t cw, uint32_t ch, uint32_t nw, uint32_t nh, uint8_t *out) {
    int rc = check_resize_args(ctx, images, n, img_w, img_h, x0, y0, cw, ch, nw, nh, out);
    if (rc) return rc;
    if (n == 0) return EMO_OK;
    EMO_CK(cudaSetDevice(ctx->device));
    return emo_launch_resize(ctx, images, n, img_w, img_h, x0, y0, cw, ch, nw, nh, out);
}

// Host pointers: images go up in slabs of ~256 MB on the copy stream while the previous slab is resized (the same two-slab
// pipeline as analyse_host), results come back per slab.
int emo_resize(emo_ctx *ctx, const uint8_t *images, uint32_t n, uint32_t img_w, uint32_t img_h, uint32_t x0, uint32_t y0, uint32_t cw,
               uint32_t ch, uint32_t nw, uint32_t nh, uint8_t *out) {
    int rc = check_resize_args(ctx, images, n, img_w, img_h, x0, y0, cw, ch, nw, nh, out);
    if (rc) return rc;
    if (n == 0) return EMO_OK;
    EMO_CK(cudaSetDevice(ctx->device));
    const size_t img_b = (size_t)img_w * img_h * 3, out_b = (size_t)nw * nh * 3;
    uint32_t slab = (uint32_t)((256ull << 20) / img_b);
    if (slab < 1) slab = 1;
    if (slab > n) slab = n;
    if ((rc = emo_ensure(ctx, &ctx->stage[0], &ctx->stage_cap[0], 2 * slab * img_b))) return rc;
    if ((rc = emo_ensure(ctx, &ctx->stage[1], &ctx->stage_cap[1], 2 * slab * out_b))) return rc;
    uint8_t *din[2] = {(uint8_t *)ctx->stage[0], (uint8_t *)ctx->stage[0] + slab * img_b};
    uint8_t *dout[2] = {(uint8_t *)ctx->stage[1], (uint8_t *)ctx->stage[1] + slab * out_b};
    uint32_t k = 0;
    for (uint32_t i0 = 0; i0 < n; i0 += slab, k++) {
        const uint32_t m = n - i0 < slab ? n - i0 : slab;
        const int b = k & 1;
        if (k >= 2) EMO_CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_pipe[2 + b], 0));
        EMO_CK(cudaMemcpyAsync(din[b], images + (size_t)i0 * img_b, m * img_b, cudaMemcpyHostToDevice, ctx->copy_stream));
        EMO_CK(cudaEventRecord(ctx->ev_pipe[b], ctx->copy_stream));
        EMO_CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_pipe[b], 0));
        if ((rc = emo_launch_resize(ctx, din[b], m, img_w, img_h, x0, y0, cw, ch, nw, nh, dout[b]))) return rc;
        EMO_CK(cudaMemcpyAsync(out + (size_t)i0 * out_b, dout[b], m * out_b, cudaMemcpyDeviceToHost, ctx->stream));
        EMO_CK(cudaEventRecord(ctx->ev_pipe[2 + b], ctx->stream));
    }
    EMO_CK(cudaStreamSynchronize(ctx->stream));
    EMO_CK(cudaStreamSynchronize(ctx->copy_stream));
    return EMO_OK;
}

}  // extern "C"

int emo_check_library_args(const void *colors, uint32_t T, uint32_t N, uint32_t ts, const void *px) {
    EMO_REQUIRE(colors, EMO_ERR_ARG, "set_library: colors is NULL");
    EMO_REQUIRE(T >= 1 && T < (1u << 30), EMO_ERR_ARG, "set_library: T=%u outside [1,2^30)", T);
    // --mode 128 (N = 16 384) is the reference's largest (main.rs:403-413); bounding N first keeps the search below in range
    EMO_REQUIRE(N >= 1 && N <= (1u << 24), EMO_ERR_ARG, "set_library: N=%u outside [1,2^24]", N);
    uint64_t dim = 1;
    while (dim * dim < N) dim++;
    EMO_REQUIRE(dim * dim == N, EMO_ERR_ARG, "set_library: N=%u is not a square", N);
    if (px) {
        EMO_REQUIRE(ts >= 1 && ts <= 4096, EMO_ERR_ARG, "set_library: tile size %u outside [1,4096]", ts);
        EMO_REQUIRE(ts % dim == 0, EMO_ERR_ARG, "Invalid tile size: Tile size must be divisible by %u", (uint32_t)dim);  // main.rs:612-615
    }
    return EMO_OK;
}

extern "C" {

}  // extern "C"

int emo_library_common(emo_ctx *ctx, uint32_t T, uint32_t N, uint32_t ts, bool has_px) {
    uint32_t dim = 1;
    while (dim * dim < N) dim++;
    uint32_t words = (3 * N + 3) / 4;
    // --mode 1..4 (N = 1, 4, 9, 16) keep the query vectors in registers; larger N (--mode 5..128, up to
    // 49 152 bytes per vector) use the tiled wide kernel with vectors padded to a multiple of 8 words.
    ctx->wide = !(words == 1 || words == 3 || words == 7 || words == 12);
    if (ctx->wide) words = (words + 7) / 8 * 8;
    if ((uint64_t)N * 3 * 255 >= (1ull << 32)) {
        emo_set_error("set_library: N=%u overflows the u32 distance", N);
        return EMO_ERR_UNSUPPORTED;
    }
    ctx->T = T; ctx->N = N; ctx->dim = dim; ctx->ts = ts; ctx->words = words;
    ctx->lut_valid = false;  // the search index belongs to the previous library
    ctx->lut16_mode = 0;
    ctx->L = (N == 1) ? T : 2 * T;  // for N == 1 a tile and its mirror coincide; the mirror can never win a tie
    uint32_t chunk = (words == 1) ? 1024 : (words == 3 ? 512 : 256);
    if (ctx->wide) chunk = 128;  // candidate tile of match_wide_kernel
    uint32_t lwin = (ctx->L + 127) / 128 * 128;  // stages are scanned in windows of 128 candidates (MATCH_WIN)
    if (lwin < chunk && !ctx->wide) chunk = lwin;
    ctx->chunk = chunk;
    ctx->n_chunks = (ctx->L + chunk - 1) / chunk;
    ctx->has_px = has_px;
    int rc;
    if ((rc = emo_ensure(ctx, (void **)&ctx->cand, &ctx->cand_cap, (size_t)ctx->n_chunks * chunk * words * 4))) return rc;
    if (has_px && (rc = emo_ensure(ctx, (void **)&ctx->lib_px, &ctx->lib_cap, (size_t)2 * T * ts * ts * 3))) return rc;
    return EMO_OK;
}

extern "C" {

int emo_set_library_dev(emo_ctx *ctx, const uint8_t *colors, const uint8_t *tile_px, uint32_t T, uint32_t N, uint32_t ts) {
    EMO_REQUIRE(ctx, EMO_ERR_ARG, "emo_set_library_dev: ctx is NULL");
    int rc = emo_check_library_args(colors, T, N, ts, tile_px);
    if (rc) return rc;
    EMO_CK(cudaSetDevice(ctx->device));
    ctx->T = 0;
    if ((rc = emo_library_common(ctx, T, N, ts, tile_px != nullptr))) { ctx->T = 0; return rc; }
    return emo_launch_build_library(ctx, colors, tile_px);
}

int emo_set_library(emo_ctx *ctx, const uint8_t *colors, const uint8_t *tile_px, uint32_t T, uint32_t N, uint32_t ts) {
    EMO_REQUIRE(ctx, EMO_ERR_ARG, "emo_set_library: ctx is NULL");
    int rc = emo_check_library_args(colors, T, N, ts, tile_px);
    if (rc) return rc;
    EMO_CK(cudaSetDevice(ctx->device));
    ctx->T = 0;
    if ((rc = emo_library_common(ctx, T, N, ts, tile_px != nullptr))) { ctx->T = 0; return rc; }
    size_t cb = (size_t)T * N * 3, pb = tile_px ? (size_t)T * ts * ts * 3 : 0;
    if ((rc = emo_ensure(ctx, &ctx->stage[0], &ctx->stage_cap[0], cb))) return rc;
    EMO_CK(cudaMemcpyAsync(ctx->stage[0], colors, cb, cudaMemcpyHostToDevice, ctx->stream));
    if (tile_px) {
        if ((rc = emo_ensure(ctx, &ctx->stage[1], &ctx->stage_cap[1], pb))) return rc;
        EMO_CK(cudaMemcpyAsync(ctx->stage[1], tile_px, pb, cudaMemcpyHostToDevice, ctx->stream));
    }
    if ((rc = emo_launch_build_library(ctx, (const uint8_t *)ctx->stage[0], tile_px ? (const uint8_t *)ctx->stage[1] : nullptr)))
        return rc;
    EMO_CK(cudaStreamSynchronize(ctx->stream));
    return EMO_OK;
}

int emo_library_info(emo_ctx *ctx, uint32_t *T, uint32_t *N, uint32_t *ts) {
    EMO_REQUIRE(ctx, EMO_ERR_ARG, "emo_library_info: ctx is NULL");
    if (T) *T = ctx->T;
    if (N) *N = ctx->T ? ctx->N : 0;
    if (ts) *ts = (ctx->T && ctx->has_px) ? ctx->ts : 0;
    return EMO_OK;
}

int emo_build_index(emo_ctx *ctx) {
    EMO_REQUIRE(ctx, EMO_ERR_ARG, "emo_build_index: ctx is NULL");
    EMO_REQUIRE(ctx->T > 0, EMO_ERR_STATE, "emo_build_index: no library set (call emo_set_library first)");
    if (!emo_index_supported(ctx)) {
        emo_set_error("emo_build_index: the colour-cube index exists for N == 1 and T <= 2^22 (N=%u, T=%u)", ctx->N, ctx->T);
        return EMO_ERR_UNSUPPORTED;
    }
    EMO_CK(cudaSetDevice(ctx->device));
    return emo_launch_build_index(ctx);
}

int emo_set_match_mode(emo_ctx *ctx, int mode) {
    EMO_REQUIRE(ctx, EMO_ERR_ARG, "emo_set_match_mode: ctx is NULL");
    EMO_REQUIRE(mode >= EMO_MATCH_AUTO && mode <= EMO_MATCH_INDEX_COMPACT, EMO_ERR_ARG,
                "emo_set_match_mode: unknown mode %d", mode);
    ctx->match_mode = mode;
    return EMO_OK;
}

static int check_match_args(emo_ctx *ctx, const void *src, uint32_t W, uint32_t H) {
    EMO_REQUIRE(ctx, EMO_ERR_ARG, "match: ctx is NULL");
    EMO_REQUIRE(ctx->T > 0, EMO_ERR_STATE, "match: no library set (call emo_set_library first)");
    EMO_REQUIRE(src, EMO_ERR_ARG, "match: src is NULL");
    EMO_REQUIRE(W > 0 && H > 0, EMO_ERR_ARG, "match: empty source %ux%u", W, H);
    // main.rs:603-611
    EMO_REQUIRE(W % ctx->dim == 0 && H % ctx->dim == 0, EMO_ERR_ARG,
                "Invalid source dimensions (%ux%u): Dimensions must be divisible by %u", W, H, ctx->dim);
    EMO_REQUIRE((uint64_t)(W / ctx->dim) * (H / ctx->dim) < (1ull << 31), EMO_ERR_ARG, "match: too many blocks");
    return EMO_OK;
}

int emo_match_dev(emo_ctx *ctx, const uint8_t *src, uint32_t W, uint32_t H, int32_t *item, uint32_t *dist) {
    int rc = check_match_args(ctx, src, W, H);
    if (rc) return rc;
    EMO_REQUIRE(item && dist, EMO_ERR_ARG, "match: item/dist is NULL");
    EMO_REQUIRE((uintptr_t)item % 4 == 0 && (uintptr_t)dist % 4 == 0, EMO_ERR_ARG, "match: item/dist must be 4-byte aligned");
    EMO_CK(cudaSetDevice(ctx->device));
    return emo_launch_match(ctx, src, W, H, item, dist);
}

int emo_match(emo_ctx *ctx, const uint8_t *src, uint32_t W, uint32_t H, int32_t *item, uint32_t *dist) {
    int rc = check_match_args(ctx, src, W, H);
    if (rc) return rc;
    EMO_REQUIRE(item && dist, EMO_ERR_ARG, "match: item/dist is NULL");
    EMO_CK(cudaSetDevice(ctx->device));
    size_t sb = (size_t)W * H * 3, Q = (size_t)(W / ctx->dim) * (H / ctx->dim);
    if ((rc = emo_ensure(ctx, &ctx->stage[2], &ctx->stage_cap[2], sb))) return rc;
    if ((rc = emo_ensure(ctx, &ctx->stage[3], &ctx->stage_cap[3], Q * 4))) return rc;
    if ((rc = emo_ensure(ctx, &ctx->stage[4], &ctx->stage_cap[4], Q * 4))) return rc;
    EMO_CK(cudaMemcpyAsync(ctx->stage[2], src, sb, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = emo_launch_match(ctx, (const uint8_t *)ctx->stage[2], W, H, (int32_t *)ctx->stage[3], (uint32_t *)ctx->stage[4])))
        return rc;
    EMO_CK(cudaMemcpyAsync(item, ctx->stage[3], Q * 4, cudaMemcpyDeviceToHost, ctx->stream));
    EMO_CK(cudaMemcpyAsync(dist, ctx->stage[4], Q * 4, cudaMemcpyDeviceToHost, ctx->stream));
    EMO_CK(cudaStreamSynchronize(ctx->stream));
    return EMO_OK;
}

// ---- ranked candidate lists (no-repeat) -----------------------------------------------------------------------------
static int check_topk_args(emo_ctx *ctx, const void *src, uint32_t W, uint32_t H, uint32_t k, const void *item, const void *dist) {
    int rc = check_match_args(ctx, src, W, H);
    if (rc) return rc;
    EMO_REQUIRE(item && dist, EMO_ERR_ARG, "topk: item/dist is NULL");
    EMO_REQUIRE(k >= 1 && k <= 1024, EMO_ERR_ARG, "topk: k=%u outside [1,1024]", k);
    if (ctx->wide) {
        emo_set_error("topk: ranked lists exist for --mode 1..4 (N = 1, 4, 9, 16), not N=%u", ctx->N);
        return EMO_ERR_UNSUPPORTED;
    }
    return EMO_OK;
}

int emo_topk_dev(emo_ctx *ctx, const uint8_t *src, uint32_t W, uint32_t H, uint32_t first, uint32_t k, const uint8_t *exclude,
                 int32_t *item, uint32_t *dist) {
    int rc = check_topk_args(ctx, src, W, H, k, item, dist);
    if (rc) return rc;
    EMO_REQUIRE((uintptr_t)item % 4 == 0 && (uintptr_t)dist % 4 == 0, EMO_ERR_ARG, "topk: item/dist must be 4-byte aligned");
    EMO_CK(cudaSetDevice(ctx->device));
    return emo_launch_topk(ctx, src, W, H, first, k, exclude, item, dist);
}

int emo_topk(emo_ctx *ctx, const uint8_t *src, uint32_t W, uint32_t H, uint32_t first, uint32_t k, const uint8_t *exclude, int32_t *item,
             uint32_t *dist) {
    int rc = check_topk_args(ctx, src, W, H, k, item, dist);
    if (rc) return rc;
    EMO_CK(cudaSetDevice(ctx->device));
    const size_t sb = (size_t)W * H * 3, ob = (size_t)(W / ctx->dim) * (H / ctx->dim) * k * 4;
    if ((rc = emo_ensure(ctx, &ctx->stage[2], &ctx->stage_cap[2], sb))) return rc;
    if ((rc = emo_ensure(ctx, &ctx->stage[3], &ctx->stage_cap[3], ob))) return rc;
    if ((rc = emo_ensure(ctx, &ctx->stage[4], &ctx->stage_cap[4], ob))) return rc;
    EMO_CK(cudaMemcpyAsync(ctx->stage[2], src, sb, cudaMemcpyHostToDevice, ctx->stream));
    const uint8_t *dex = nullptr;
    if (exclude) {
        if ((rc = emo_ensure(ctx, &ctx->stage[5], &ctx->stage_cap[5], ctx->T))) return rc;
        EMO_CK(cudaMemcpyAsync(ctx->stage[5], exclude, ctx->T, cudaMemcpyHostToDevice, ctx->stream));
        dex = (const uint8_t *)ctx->stage[5];
    }
    if ((rc = emo_launch_topk(ctx, (const uint8_t *)ctx->stage[2], W, H, first, k, dex, (int32_t *)ctx->stage[3],
                              (uint32_t *)ctx->stage[4])))
        return rc;
    EMO_CK(cudaMemcpyAsync(item, ctx->stage[3], ob, cudaMemcpyDeviceToHost, ctx->stream));
    EMO_CK(cudaMemcpyAsync(dist, ctx->stage[4], ob, cudaMemcpyDeviceToHost, ctx->stream));
    EMO_CK(cudaStreamSynchronize(ctx->stream));
    return EMO_OK;
}

static int check_compose_args(emo_ctx *ctx, const void *item, const void *src, uint32_t W, uint32_t H, uint32_t oc,
                              const void *out) {
    EMO_REQUIRE(ctx, EMO_ERR_ARG, "compose: ctx is NULL");
    EMO_REQUIRE(ctx->T > 0 && ctx->has_px, EMO_ERR_STATE, "compose: no tile pixels resident (emo_set_library with tile_px)");
    EMO_REQUIRE(item && out, EMO_ERR_ARG, "compose: item/out is NULL");
    EMO_REQUIRE(oc == 3 || oc == 4, EMO_ERR_ARG, "compose: out_channels must be 3 (RGB) or 4 (RGBA + tint), got %u", oc);
    EMO_REQUIRE(oc == 3 || src, EMO_ERR_ARG, "compose: tint needs the source image");
    EMO_REQUIRE(W > 0 && H > 0 && W % ctx->dim == 0 && H % ctx->dim == 0, EMO_ERR_ARG,
                "Invalid source dimensions (%ux%u): Dimensions must be divisible by %u", W, H, ctx->dim);
    return EMO_OK;
}

int emo_compose_dev(emo_ctx *ctx, const int32_t *item, const uint8_t *src, uint32_t W, uint32_t H, uint32_t oc,
                    uint8_t tint_alpha, uint8_t *out) {
    int rc = check_compose_args(ctx, item, src, W, H, oc, out);
    if (rc) return rc;
    EMO_REQUIRE((uintptr_t)item % 4 == 0, EMO_ERR_ARG, "compose: item map must be 4-byte aligned");
    EMO_REQUIRE(oc == 3 || (uintptr_t)out % 4 == 0, EMO_ERR_ARG, "compose: RGBA output must be 4-byte aligned");
    EMO_CK(cudaSetDevice(ctx->device));
    return emo_launch_compose(ctx, item, src, W, H, oc, tint_alpha, out);
}

int emo_compose(emo_ctx *ctx, const int32_t *item, const uint8_t *src, uint32_t W, uint32_t H, uint32_t oc,
                uint8_t tint_alpha, uint8_t *out) {
    int rc = check_compose_args(ctx, item, src, W, H, oc, out);
    if (rc) return rc;
    EMO_CK(cudaSetDevice(ctx->device));
    uint32_t bw = W / ctx->dim, bh = H / ctx->dim;
    size_t Q = (size_t)bw * bh, ob = Q * ctx->ts * ctx->ts * oc, sb = (size_t)W * H * 3;
    if ((rc = emo_ensure(ctx, &ctx->stage[3], &ctx->stage_cap[3], Q * 4))) return rc;
    if ((rc = emo_ensure(ctx, &ctx->stage[5], &ctx->stage_cap[5], ob))) return rc;
    EMO_CK(cudaMemcpyAsync(ctx->stage[3], item, Q * 4, cudaMemcpyHostToDevice, ctx->stream));
    const uint8_t *dsrc = nullptr;
    if (oc == 4) {
        if ((rc = emo_ensure(ctx, &ctx->stage[2], &ctx->stage_cap[2], sb))) return rc;
        EMO_CK(cudaMemcpyAsync(ctx->stage[2], src, sb, cudaMemcpyHostToDevice, ctx->stream));
        dsrc = (const uint8_t *)ctx->stage[2];
    }
    if ((rc = emo_launch_compose(ctx, (const int32_t *)ctx->stage[3], dsrc, W, H, oc, tint_alpha, (uint8_t *)ctx->stage[5])))
        return rc;
    EMO_CK(cudaMemcpyAsync(out, ctx->stage[5], ob, cudaMemcpyDeviceToHost, ctx->stream));
    EMO_CK(cudaStreamSynchronize(ctx->stream));
    return emo_check_device_flag(ctx);
}

int emo_compose_overlay_dev(emo_ctx *ctx, const int32_t *item, uint32_t W, uint32_t H, const uint8_t *overlay, uint32_t ow,
                            uint32_t oh, uint8_t tint_alpha, uint8_t *out) {
    int rc = check_compose_args(ctx, item, overlay, W, H, 4, out);
    if (rc) return rc;
    EMO_REQUIRE(ow > 0 && oh > 0, EMO_ERR_ARG, "compose_overlay: empty overlay %ux%u", ow, oh);
    EMO_REQUIRE((uintptr_t)item % 4 == 0 && (uintptr_t)out % 4 == 0, EMO_ERR_ARG, "compose_overlay: item/out must be 4-byte aligned");
    EMO_CK(cudaSetDevice(ctx->device));
    return emo_launch_compose_overlay(ctx, item, W, H, overlay, ow, oh, tint_alpha, out);
}

int emo_compose_overlay(emo_ctx *ctx, const int32_t *item, uint32_t W, uint32_t H, const uint8_t *overlay, uint32_t ow, uint32_t oh,
                        uint8_t tint_alpha, uint8_t *out) {
    int rc = check_compose_args(ctx, item, overlay, W, H, 4, out);
    if (rc) return rc;
    EMO_REQUIRE(ow > 0 && oh > 0, EMO_ERR_ARG, "compose_overlay: empty overlay %ux%u", ow, oh);
    EMO_CK(cudaSetDevice(ctx->device));
    const uint32_t bw = W / ctx->dim, bh = H / ctx->dim;
    const size_t Q = (size_t)bw * bh, ob = Q * ctx->ts * ctx->ts * 4, sb = (size_t)ow * oh * 3;
    if ((rc = emo_ensure(ctx, &ctx->stage[3], &ctx->stage_cap[3], Q * 4))) return rc;
    if ((rc = emo_ensure(ctx, &ctx->stage[5], &ctx->stage_cap[5], ob))) return rc;
    if ((rc = emo_ensure(ctx, &ctx->stage[2], &ctx->stage_cap[2], sb))) return rc;
    EMO_CK(cudaMemcpyAsync(ctx->stage[3], item, Q * 4, cudaMemcpyHostToDevice, ctx->stream));
    EMO_CK(cudaMemcpyAsync(ctx->stage[2], overlay, sb, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = emo_launch_compose_overlay(ctx, (const int32_t *)ctx->stage[3], W, H, (const uint8_t *)ctx->stage[2], ow, oh, tint_alpha,
                                         (uint8_t *)ctx->stage[5])))
        return rc;
    EMO_CK(cudaMemcpyAsync(out, ctx->stage[5], ob, cudaMemcpyDeviceToHost, ctx->stream));
    EMO_CK(cudaStreamSynchronize(ctx->stream));
    return emo_check_device_flag(ctx);
}

// Whole path with host buffers: the source goes up once, then block-row chunks are matched and
// composed on the compute stream while the previous chunk's output drains to the host on the
// copy stream (two device output buffers).
// match + compose of one (stripe of an) image, device pointers, back to back on the ctx stream
static int mosaic_launch(emo_ctx *ctx, const uint8_t *src, uint32_t W, uint32_t H, uint32_t oc, uint8_t tint_alpha, int32_t *item,
                         uint32_t *dist, uint8_t *out) {
    int rc = emo_launch_match(ctx, src, W, H, item, dist);
    if (rc) return rc;
    return emo_launch_compose(ctx, item, src, W, H, oc, tint_alpha, out);
}

int emo_mosaic_dev(emo_ctx *ctx, const uint8_t *src, uint32_t W, uint32_t H, uint32_t oc, uint8_t tint_alpha, int32_t *item,
                   uint32_t *dist, uint8_t *out) {
    int rc = check_match_args(ctx, src, W, H);
    if (rc) return rc;
    EMO_REQUIRE(item && dist, EMO_ERR_ARG, "mosaic: item/dist is NULL");
    EMO_REQUIRE((uintptr_t)item % 4 == 0 && (uintptr_t)dist % 4 == 0, EMO_ERR_ARG, "mosaic: item/dist must be 4-byte aligned");
    if ((rc = check_compose_args(ctx, item, src, W, H, oc, out))) return rc;
    EMO_REQUIRE(oc == 3 || (uintptr_t)out % 4 == 0, EMO_ERR_ARG, "mosaic: RGBA output must be 4-byte aligned");
    EMO_CK(cudaSetDevice(ctx->device));
    if ((rc = emo_prepare_match(ctx, (uint64_t)(W / ctx->dim) * (H / ctx->dim)))) return rc;
    // (Replaying a captured graph of the two launches was measured and dropped: 592.3 vs 590.0 us per C4 step and 80.7 vs 78.7 us on
    // a 512-row stripe — the programmatic-dependent launches on the stream already overlap the launch gaps)
    return mosaic_launch(ctx, src, W, H, oc, tint_alpha, item, dist, out);
}

int emo_mosaic(emo_ctx *ctx, const uint8_t *src, uint32_t W, uint32_t H, uint32_t oc, uint8_t tint_alpha,
               int32_t *item, uint32_t *dist, uint8_t *out) {
    return emo_mosaic_host_impl(ctx, src, W, H, oc, tint_alpha, item, dist, out, 0);
}

}  // extern "C"

// Block rows per pipeline step of emo_mosaic.  A chunk is ~1/16 of the output, between 8 and 64 MB: long enough copies to run
// at the PCIe rate, short enough that the first chunk's kernels and the last chunk's copy — the two ends of the pipeline that
// overlap with nothing — stay small next to the whole (C2's 201 MB: 16 chunks of 12.6 MB instead of 3 of 67 MB); but at least
// ~1.2 M queries per launch when that still leaves >= 8 chunks (C4: 293 block rows = 230 MB); at least one block row.
static uint32_t mosaic_rows_per_chunk(uint32_t bw, uint32_t bh, size_t row_out) {
    if (!row_out) row_out = 1;
    size_t target = row_out * bh / 16;
    if (target < (8ull << 20)) target = 8ull << 20;
    if (target > (64ull << 20)) target = 64ull << 20;
    uint32_t rows = (uint32_t)(target / row_out);
    const uint32_t rows_for_q = ((1200000u + bw - 1) / bw);
    if (rows < rows_for_q && rows_for_q * 8 <= bh) rows = rows_for_q;
    if (rows < 1) rows = 1;
    if (rows > bh) rows = bh;
    return rows;
}

extern "C" int emo_reserve(emo_ctx *ctx, uint32_t W, uint32_t H, uint32_t oc) {
    EMO_REQUIRE(ctx, EMO_ERR_ARG, "emo_reserve: ctx is NULL");
    EMO_REQUIRE(ctx->T > 0, EMO_ERR_STATE, "emo_reserve: no library set (the staging sizes depend on its cell grid and tile size)");
    EMO_REQUIRE(W > 0 && H > 0 && (oc == 3 || oc == 4), EMO_ERR_ARG, "emo_reserve: bad geometry %ux%u, %u channels", W, H, oc);
    EMO_CK(cudaSetDevice(ctx->device));
    const uint32_t dim = ctx->dim, bw = (W + dim - 1) / dim, bh = (H + dim - 1) / dim;
    const size_t Q = (size_t)bw * bh, row_out = (size_t)bw * ctx->ts * ctx->ts * oc;
    int rc;
    if ((rc = emo_ensure(ctx, &ctx->stage[2], &ctx->stage_cap[2], (size_t)bw * dim * bh * dim * 3))) return rc;
    if ((rc = emo_ensure(ctx, &ctx->stage[3], &ctx->stage_cap[3], Q * 4))) return rc;
    if ((rc = emo_ensure(ctx, &ctx->stage[4], &ctx->stage_cap[4], Q * 4))) return rc;
    if (ctx->has_px && (rc = emo_ensure(ctx, &ctx->stage[5], &ctx->stage_cap[5], row_out * mosaic_rows_per_chunk(bw, bh, row_out) * 2))) return rc;
    if (emo_index_supported(ctx) && ctx->match_mode != EMO_MATCH_SCAN && (rc = emo_index_reserve(ctx))) return rc;  // 1to1: the 64 + 32 MiB tables
    if (!ctx->wide && (rc = emo_ensure(ctx, (void **)&ctx->keys, &ctx->keys_cap, Q * 8))) return rc;  // merge keys of a split scan
    if (oc == 4 && ctx->tint.alpha < 0 && (rc = emo_prepare_tint(ctx, 127))) return rc;              // the blend tables (rebuilt per alpha, allocated once)
    return EMO_OK;
}

// total_queries: the block count the 1to1 index rule is applied to (0 = this image's own); a multi-GPU caller passes the
// whole image's count so that every stripe takes the same decision as a single-GPU run would.
int emo_mosaic_host_impl(emo_ctx *ctx, const uint8_t *src, uint32_t W, uint32_t H, uint32_t oc, uint8_t tint_alpha, int32_t *item,
                         uint32_t *dist, uint8_t *out, uint64_t total_queries) {
    int rc = check_match_args(ctx, src, W, H);
    if (rc) return rc;
    EMO_REQUIRE(ctx->has_px, EMO_ERR_STATE, "mosaic: no tile pixels resident (emo_set_library with tile_px)");
    EMO_REQUIRE(out, EMO_ERR_ARG, "mosaic: out is NULL");
    EMO_REQUIRE(oc == 3 || oc == 4, EMO_ERR_ARG, "mosaic: out_channels must be 3 or 4, got %u", oc);
    EMO_CK(cudaSetDevice(ctx->device));
    const uint32_t dim = ctx->dim, ts = ctx->ts, bw = W / dim, bh = H / dim;
    const size_t Q = (size_t)bw * bh, sb = (size_t)W * H * 3;
    const size_t row_out = (size_t)bw * ts * ts * oc;  // output bytes per block row
    const uint32_t rows_per_chunk = mosaic_rows_per_chunk(bw, bh, row_out);
    const size_t chunk_out = row_out * rows_per_chunk;
    if ((rc = emo_ensure(ctx, &ctx->stage[2], &ctx->stage_cap[2], sb))) return rc;
    if ((rc = emo_ensure(ctx, &ctx->stage[3], &ctx->stage_cap[3], Q * 4))) return rc;
    if ((rc = emo_ensure(ctx, &ctx->stage[4], &ctx->stage_cap[4], Q * 4))) return rc;
    if ((rc = emo_ensure(ctx, &ctx->stage[5], &ctx->stage_cap[5], chunk_out * 2))) return rc;
    uint8_t *dsrc = (uint8_t *)ctx->stage[2];
    int32_t *ditem = (int32_t *)ctx->stage[3];
    uint32_t *ddist = (uint32_t *)ctx->stage[4];
    uint8_t *dout[2] = {(uint8_t *)ctx->stage[5], (uint8_t *)ctx->stage[5] + chunk_out};
    EMO_CK(cudaMemcpyAsync(dsrc, src, sb, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = emo_prepare_match(ctx, total_queries ? total_queries : Q))) return rc;  // decide on the 1to1 index for the whole image, not per chunk
    uint32_t k = 0;
    for (uint32_t r0 = 0; r0 < bh; r0 += rows_per_chunk, k++) {
        uint32_t nr = bh - r0 < rows_per_chunk ? bh - r0 : rows_per_chunk;
        const uint8_t *s = dsrc + (size_t)r0 * dim * W * 3;
        int b = k & 1;
        // the buffer must have drained (chunk k-2) before it is overwritten
        if (k >= 2) EMO_CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_pipe[2 + b], 0));
        if ((rc = mosaic_launch(ctx, s, W, nr * dim, oc, tint_alpha, ditem + (size_t)r0 * bw, ddist + (size_t)r0 * bw, dout[b])))
            return rc;
        EMO_CK(cudaEventRecord(ctx->ev_pipe[b], ctx->stream));
        EMO_CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_pipe[b], 0));
        EMO_CK(cudaMemcpyAsync(out + (size_t)r0 * row_out, dout[b], (size_t)nr * row_out, cudaMemcpyDeviceToHost,
                               ctx->copy_stream));
        EMO_CK(cudaEventRecord(ctx->ev_pipe[2 + b], ctx->copy_stream));
    }
    if (item) EMO_CK(cudaMemcpyAsync(item, ditem, Q * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (dist) EMO_CK(cudaMemcpyAsync(dist, ddist, Q * 4, cudaMemcpyDeviceToHost, ctx->stream));
    EMO_CK(cudaStreamSynchronize(ctx->stream));
    EMO_CK(cudaStreamSynchronize(ctx->copy_stream));
    return emo_check_device_flag(ctx);
}
