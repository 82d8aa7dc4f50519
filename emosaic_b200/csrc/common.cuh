// common.cuh — context, error plumbing and sm_100a PTX helpers shared by the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/emosaic_cuda.h"

// ---------------------------------------------------------------------------------------
// error plumbing (nothing unwinds across the C ABI)
// ---------------------------------------------------------------------------------------
void emo_set_error(const char *fmt, ...);

#define EMO_CK(call)                                                                              \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            emo_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return e__ == cudaErrorMemoryAllocation ? EMO_ERR_OOM : EMO_ERR_CUDA;                 \
        }                                                                                         \
    } while (0)

#define EMO_REQUIRE(cond, code, ...)                                                              \
    do {                                                                                          \
        if (!(cond)) {                                                                            \
            emo_set_error(__VA_ARGS__);                                                           \
            return (code);                                                                        \
        }                                                                                         \
    } while (0)

#define EMO_LAUNCH_CHECK(ctx)                                                                     \
    do {                                                                                          \
        (ctx)->launches++;                                                                        \
        EMO_CK(cudaGetLastError());                                                               \
    } while (0)

// ---------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------
// 1to1 search index keys (index.cu): dist << 22 | tile
static constexpr uint32_t EMO_IDX_TILE_BITS = 22;
static constexpr uint32_t EMO_IDX_TILE_MASK = (1u << EMO_IDX_TILE_BITS) - 1;

struct emo_tint_tables {
    int alpha = -1;           // alpha the tables were built for (-1: none)
    int K = 0;                // max exceptional bg values per fg (see compose.cu)
    uint8_t alpha_out = 255;  // trunc(255 * alpha_final)
    uint8_t *lut = nullptr;   // [256 bg][256 fg] exact f32 blend results          (device)
    uint8_t *excv = nullptr;  // [4][256 fg] exceptional bg values                   (device)
    uint8_t *excm = nullptr;  // [4][256 fg] 0x80 where excv is valid                (device)
    uint8_t *cadd = nullptr;  // [256 fg] additive constant 0/1 of the lane formula      (device)
    uint32_t *meta = nullptr; // [0] = max exceptions per fg, [1] = alpha byte, [2] = sanity errors, [3] = non-uniform fg (device)
};

struct emo_resize_state;  // resize.cu: cached tap tables + the f32 intermediate image

struct emo_ctx {
    int device = 0;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;  // stream all work is issued on (own or external)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    cudaEvent_t ev_pipe[4] = {nullptr, nullptr, nullptr, nullptr};
    uint64_t launches = 0;
    std::vector<cudaEvent_t> marks;

    // resident library
    uint32_t T = 0, N = 0, dim = 0, ts = 0;
    uint32_t words = 0;       // packed u32 words per candidate vector = ceil(3N/4)
    bool wide = false;        // N outside {1,4,9,16}: vectors padded to a multiple of 8 words, match_wide_kernel
    uint32_t *qvec = nullptr; // wide path scratch: packed query vectors [Qpad][words]
    size_t qvec_cap = 0;
    uint32_t L = 0;           // real candidates (T for N==1, 2T otherwise)
    uint32_t chunk = 0;       // candidates per shared-memory stage
    uint32_t n_chunks = 0;    // Lpad / chunk
    uint32_t *cand = nullptr; // [n_chunks*chunk][words]
    size_t cand_cap = 0;
    uint8_t *lib_px = nullptr; // [2T][ts][ts][3]: entry 2t = tile t, 2t+1 = tile t mirrored
    size_t lib_cap = 0;
    bool has_px = false;
    uint32_t *lut = nullptr;  // 1to1 search index: [256 b][256 g][256 r] keys dist << 22 | tile (index.cu), 64 MiB
    bool lut_valid = false;   // built for the resident library
    uint32_t *idx_seeded = nullptr;  // one bit per cell: a library colour lives here (2 MiB; index.cu, build only)
    uint16_t *lut16 = nullptr;  // compact form: slot of the winner per cell, 32 MiB (index.cu)
    int lut16_mode = 0;         // 0: none, 1: slot = tile index (T <= 65 536), 2: slot -> idx_entry {tile, colour}, 3: 2 pending the winner count
    uint32_t lut16_slots = 0;
    uint32_t *idx_slot_of_tile = nullptr;  // build scratch of mode 2
    size_t idx_slot_cap = 0;
    uint2 *idx_entry = nullptr;            // [65 536] {tile, colour} + the winner counter
    volatile uint32_t *idx_count_host = nullptr;  // pinned: the winner count of the index being built
    cudaEvent_t idx_count_ev = nullptr;
    int match_mode = 0;       // EMO_MATCH_AUTO / SCAN / INDEX

    // scratch
    unsigned long long *keys = nullptr;  // [Q] packed (dist<<32 | candidate) for split matching
    size_t keys_cap = 0;
    int *err_flag = nullptr;             // device: bit0 = item out of range in compose
    void *stage[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // device staging for the host API
    size_t stage_cap[6] = {0, 0, 0, 0, 0, 0};

    emo_tint_tables tint;
    emo_resize_state *resize = nullptr;

    // multi-GPU (comm.cu): the ctx's NCCL communicator (ncclComm_t), its rank, and a 64-byte device scratch for headers
    void *comm = nullptr;
    int rank = 0, world = 1;
    bool comm_owned = false;  // false: the communicator belongs to an emo_group
    uint32_t *comm_hdr = nullptr;
};

int emo_ensure(emo_ctx *ctx, void **p, size_t *cap, size_t bytes);  // grow-only device buffer
void emo_comm_release(emo_ctx *ctx);                                // comm.cu: drops the ctx's communicator (emo_destroy)
// argument checks and library bookkeeping of ctx.cu, shared with comm.cu
int emo_check_analyse_args(const void *tiles, uint64_t T, uint32_t ts, uint32_t dim, const void *out);
int emo_check_library_args(const void *colors, uint32_t T, uint32_t N, uint32_t ts, const void *px);
int emo_library_common(emo_ctx *ctx, uint32_t T, uint32_t N, uint32_t ts, bool has_px);
// host-pointer pipelines of ctx.cu, shared with the multi-GPU group calls of comm.cu
int emo_analyse_host_impl(emo_ctx *ctx, const uint8_t *tiles, uint64_t T, uint32_t ts, uint32_t dim, uint8_t *out1, uint8_t *out4,
                          bool fused);
int emo_mosaic_host_impl(emo_ctx *ctx, const uint8_t *src, uint32_t W, uint32_t H, uint32_t oc, uint8_t tint_alpha, int32_t *item,
                         uint32_t *dist, uint8_t *out, uint64_t total_queries);
int emo_check_device_flag(emo_ctx *ctx);

// kernels' host launchers (device pointers, async on ctx->stream)
int emo_launch_analyse(emo_ctx *ctx, const uint8_t *tiles, uint64_t T, uint32_t ts, uint32_t dim, uint8_t *out);
int emo_launch_analyse_fused(emo_ctx *ctx, const uint8_t *tiles, uint64_t T, uint32_t ts, uint8_t *out1, uint8_t *out4);
int emo_launch_resize(emo_ctx *ctx, const uint8_t *images, uint32_t n, uint32_t img_w, uint32_t img_h, uint32_t x0, uint32_t y0,
                      uint32_t cw, uint32_t ch, uint32_t nw, uint32_t nh, uint8_t *out);
void emo_resize_state_free(emo_resize_state *s);
int emo_launch_build_library(emo_ctx *ctx, const uint8_t *colors, const uint8_t *tile_px);
int emo_launch_match(emo_ctx *ctx, const uint8_t *src, uint32_t W, uint32_t H, int32_t *item, uint32_t *dist);
int emo_launch_topk(emo_ctx *ctx, const uint8_t *src, uint32_t W, uint32_t H, uint32_t first, uint32_t k, const uint8_t *exclude,
                    int32_t *item, uint32_t *dist);
int emo_launch_build_index(emo_ctx *ctx);
int emo_index_reserve(emo_ctx *ctx);  // allocates the index tables without building them
int emo_prepare_match(emo_ctx *ctx, uint64_t queries);  // builds the 1to1 index when the mode / size rule asks for it
bool emo_index_supported(const emo_ctx *ctx);
int emo_launch_match_index(emo_ctx *ctx, const uint8_t *src, uint32_t W, uint32_t H, int32_t *item, uint32_t *dist);
int emo_launch_compose(emo_ctx *ctx, const int32_t *item, const uint8_t *src, uint32_t W, uint32_t H,
                       uint32_t out_channels, uint8_t tint_alpha, uint8_t *out);
int emo_prepare_tint(emo_ctx *ctx, uint8_t alpha);
int emo_launch_compose_overlay(emo_ctx *ctx, const int32_t *item, uint32_t W, uint32_t H, const uint8_t *overlay, uint32_t ow,
                               uint32_t oh, uint8_t tint_alpha, uint8_t *out);

// ---------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------
#ifdef __CUDACC__
// launch with programmatic stream serialization (see grid_dependency_wait)
template <typename... KArgs, typename... Args>
static inline cudaError_t emo_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                         Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// Same wait for a thread that is expected to wait long (the producer on an `empty` barrier): the try_wait
// carries a suspend-time hint so the warp sleeps in hardware instead of spinning through issue slots.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAITR_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONER_%=;\n\t"
        "bra WAITR_%=;\n\t"
        "DONER_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(20000u)
        : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// TMA 1-D bulk copy shared -> global (bulk async-group completion)
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
// Programmatic dependent launch: the kernel may be scheduled while the previous kernel of the stream drains; it must
// not touch that kernel's results before grid_dependency_wait() (a no-op for a normally launched kernel).
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// L2 eviction-priority policies (createpolicy) and the loads / stores that carry them
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint32_t ldg_nc_hint_u32(const void *p, uint64_t pol) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ uint32_t ldg_nc_hint_u16(const void *p, uint64_t pol) {
    uint16_t r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u16 %0, [%1], %2;" : "=h"(r) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void stg_hint_v4(void *p, uint4 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g_hint(void *dst_gmem, const void *src_smem, uint32_t bytes, uint64_t pol) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst_gmem),
                 "r"(smem_u32(src_smem)), "r"(bytes), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit_wait_read() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ uint32_t sad4(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;  // SASS: VABSDIFF4.U8.ACC — c + sum over the 4 bytes of |a_i - b_i|
}

__device__ __forceinline__ uint4 ldg_nc_v4(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_cs_v4(void *p, uint4 v) {
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void stg_cs_v2(void *p, uint2 v) {
    asm volatile("st.global.cs.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void stg_cs_u32(void *p, uint32_t v) {
    asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
#endif
