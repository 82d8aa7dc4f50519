// comm.cu — the multi-GPU half of the C ABI (include/emosaic_cuda.h, "multi-GPU").
//
// What is sharded are the reference's own parallel tasks: one block row of the source per rayon task in render()
// ((0..H).into_par_iter().step_by(step), src/mosaic/rendering.rs:68-89, strips merged at :91-99) and one tile per task in
// the analysis build (src/main.rs:760-794).  GPU r of n takes a contiguous range of them; the library is replicated once
// with ncclBroadcast over NVLink / NVSwitch; output stripes are disjoint row ranges of one host image and every GPU copies
// its stripe straight to its row offset.  There is no collective inside the match / compose loop.
//
// NCCL is resolved with dlopen at the first call that needs it: a process that already carries a libnccl.so.2 (PyTorch
// bundles its own) keeps exactly one copy, and single-GPU hosts need none.
#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only; nothing here links against libnccl
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <thread>

#include "common.cuh"

// ---------------------------------------------------------------------------------------
// run-time binding of NCCL
// ---------------------------------------------------------------------------------------
namespace {
struct NcclApi {
    void *handle = nullptr;
    std::string origin, error;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
};
NcclApi g_nccl;
std::once_flag g_nccl_once;

void nccl_bind() {
    NcclApi &a = g_nccl;
    if (const char *p = getenv("EMO_NCCL_LIB")) {
        a.handle = dlopen(p, RTLD_NOW | RTLD_GLOBAL);
        a.origin = p;
        if (!a.handle) { a.error = dlerror(); return; }
    }
    if (!a.handle) {  // a copy the process already loaded (e.g. the one PyTorch ships): never load a second NCCL next to it
        a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        a.origin = "libnccl.so.2 (already in the process)";
    }
    if (!a.handle) {
        a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        a.origin = "libnccl.so.2";
    }
    if (!a.handle) {
        a.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        a.origin = "libnccl.so";
    }
    if (!a.handle) {
        a.error = "libnccl.so.2 not found (set EMO_NCCL_LIB to its path)";
        return;
    }
#define EMO_BIND(name)                                                          \
    a.name = reinterpret_cast<decltype(a.name)>(dlsym(a.handle, "nccl" #name)); \
    if (!a.name) { a.error = "symbol nccl" #name " missing in " + a.origin; return; }
    EMO_BIND(GetVersion) EMO_BIND(GetUniqueId) EMO_BIND(CommInitRank) EMO_BIND(CommInitAll) EMO_BIND(CommDestroy)
    EMO_BIND(GetErrorString) EMO_BIND(Broadcast) EMO_BIND(AllGather) EMO_BIND(GroupStart) EMO_BIND(GroupEnd)
#undef EMO_BIND
}

// nullptr + emo_last_error when NCCL cannot be used
const NcclApi *nccl() {
    std::call_once(g_nccl_once, nccl_bind);
    if (!g_nccl.error.empty()) {
        emo_set_error("NCCL is not available: %s", g_nccl.error.c_str());
        return nullptr;
    }
    return &g_nccl;
}
}  // namespace

#define EMO_NCCL(api, call)                                                                                   \
    do {                                                                                                      \
        ncclResult_t r__ = (call);                                                                            \
        if (r__ != ncclSuccess) {                                                                             \
            emo_set_error("%s failed: %s (%s:%d)", #call, (api)->GetErrorString(r__), __FILE__, __LINE__);    \
            return EMO_ERR_NCCL;                                                                              \
        }                                                                                                     \
    } while (0)

void emo_comm_release(emo_ctx *ctx) {
    if (ctx->comm && ctx->comm_owned && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)ctx->comm);
    ctx->comm = nullptr;
    ctx->comm_owned = false;
    ctx->rank = 0;
    ctx->world = 1;
}

extern "C" {

void emo_stripe_bounds(uint64_t units, int world, int rank, uint64_t *start, uint64_t *stop) {
    uint64_t a = 0, b = 0;
    if (world >= 1 && rank >= 0 && rank < world) {
        const uint64_t base = units / (uint64_t)world, extra = units % (uint64_t)world, r = (uint64_t)rank;
        a = r * base + (r < extra ? r : extra);
        b = a + base + (r < extra ? 1 : 0);
    }
    if (start) *start = a;
    if (stop) *stop = b;
}

int emo_host_register(void *p, size_t bytes) {
    EMO_REQUIRE(p && bytes, EMO_ERR_ARG, "emo_host_register: empty range");
    EMO_CK(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return EMO_OK;
}
int emo_host_unregister(void *p) {
    EMO_REQUIRE(p, EMO_ERR_ARG, "emo_host_unregister: NULL");
    EMO_CK(cudaHostUnregister(p));
    return EMO_OK;
}

// ---------------------------------------------------------------------------------------
// (a) one process per GPU
// ---------------------------------------------------------------------------------------
int emo_comm_unique_id(void *id_out) {
    EMO_REQUIRE(id_out, EMO_ERR_ARG, "emo_comm_unique_id: id_out is NULL");
    static_assert(EMO_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "id size");
    const NcclApi *api = nccl();
    if (!api) return EMO_ERR_NCCL;
    ncclUniqueId id;
    EMO_NCCL(api, api->GetUniqueId(&id));
    memcpy(id_out, &id, sizeof id);
    return EMO_OK;
}

int emo_comm_init_rank(emo_ctx *ctx, const void *id, int rank, int world) {
    EMO_REQUIRE(ctx && id, EMO_ERR_ARG, "emo_comm_init_rank: NULL argument");
    EMO_REQUIRE(world >= 1 && rank >= 0 && rank < world, EMO_ERR_ARG, "emo_comm_init_rank: rank %d outside [0,%d)", rank, world);
    EMO_REQUIRE(!ctx->comm, EMO_ERR_STATE, "emo_comm_init_rank: the ctx already belongs to a communicator");
    const NcclApi *api = nccl();
    if (!api) return EMO_ERR_NCCL;
    EMO_CK(cudaSetDevice(ctx->device));
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof uid);
    ncclComm_t comm = nullptr;
    EMO_NCCL(api, api->CommInitRank(&comm, world, uid, rank));
    ctx->comm = comm;
    ctx->comm_owned = true;
    ctx->rank = rank;
    ctx->world = world;
    return EMO_OK;
}

int emo_comm_info(emo_ctx *ctx, int *rank, int *world, int *nccl_version) {
    EMO_REQUIRE(ctx, EMO_ERR_ARG, "emo_comm_info: ctx is NULL");
    if (rank) *rank = ctx->rank;
    if (world) *world = ctx->world;
    if (nccl_version) {
        *nccl_version = 0;
        if (ctx->comm) {
            const NcclApi *api = nccl();
            if (!api) return EMO_ERR_NCCL;
            EMO_NCCL(api, api->GetVersion(nccl_version));
        }
    }
    return EMO_OK;
}

int emo_comm_broadcast_dev(emo_ctx *ctx, void *buf, size_t bytes, int root) {
    EMO_REQUIRE(ctx && (buf || bytes == 0), EMO_ERR_ARG, "emo_comm_broadcast_dev: NULL argument");
    EMO_REQUIRE(root >= 0 && root < ctx->world, EMO_ERR_ARG, "emo_comm_broadcast_dev: root %d outside [0,%d)", root, ctx->world);
    if (ctx->world == 1 || bytes == 0) return EMO_OK;
    const NcclApi *api = nccl();
    if (!api) return EMO_ERR_NCCL;
    EMO_CK(cudaSetDevice(ctx->device));
    EMO_NCCL(api, api->Broadcast(buf, buf, bytes, ncclUint8, root, (ncclComm_t)ctx->comm, ctx->stream));
    return EMO_OK;
}

int emo_comm_allgather_analysis_dev(emo_ctx *ctx, const uint8_t *local, uint64_t T, uint32_t bpt, uint8_t *all) {
    EMO_REQUIRE(ctx && all && bpt, EMO_ERR_ARG, "emo_comm_allgather_analysis_dev: NULL argument");
    uint64_t a, b;
    emo_stripe_bounds(T, ctx->world, ctx->rank, &a, &b);
    EMO_REQUIRE(local || a == b, EMO_ERR_ARG, "emo_comm_allgather_analysis_dev: local is NULL");
    EMO_CK(cudaSetDevice(ctx->device));
    if (ctx->world == 1) {
        if (local != all && T) EMO_CK(cudaMemcpyAsync(all, local, T * bpt, cudaMemcpyDeviceToDevice, ctx->stream));
        return EMO_OK;
    }
    const NcclApi *api = nccl();
    if (!api) return EMO_ERR_NCCL;
    ncclComm_t comm = (ncclComm_t)ctx->comm;
    if (T % (uint64_t)ctx->world == 0) {  // equal shards: one all-gather
        EMO_NCCL(api, api->AllGather(local, all, (b - a) * bpt, ncclUint8, comm, ctx->stream));
        return EMO_OK;
    }
    // ragged shards (sizes differ by one tile): the all-gather-v is one fused group of broadcasts, rank r the root of its range
    EMO_NCCL(api, api->GroupStart());
    for (int r = 0; r < ctx->world; r++) {
        uint64_t ra, rb;
        emo_stripe_bounds(T, ctx->world, r, &ra, &rb);
        if (rb == ra) continue;
        const ncclResult_t res = api->Broadcast(r == ctx->rank ? (const void *)local : (const void *)(all + ra * bpt), all + ra * bpt,
                                                (rb - ra) * bpt, ncclUint8, r, comm, ctx->stream);
        if (res != ncclSuccess) {
            api->GroupEnd();
            emo_set_error("ncclBroadcast failed: %s", api->GetErrorString(res));
            return EMO_ERR_NCCL;
        }
    }
    EMO_NCCL(api, api->GroupEnd());
    return EMO_OK;
}

static int comm_set_library(emo_ctx *ctx, const uint8_t *colors, const uint8_t *tile_px, uint32_t T, uint32_t N, uint32_t ts, int root,
                            bool dev_ptrs) {
    EMO_REQUIRE(ctx, EMO_ERR_ARG, "emo_comm_set_library: ctx is NULL");
    EMO_REQUIRE(root >= 0 && root < ctx->world, EMO_ERR_ARG, "emo_comm_set_library: root %d outside [0,%d)", root, ctx->world);
    if (ctx->world == 1)
        return dev_ptrs ? emo_set_library_dev(ctx, colors, tile_px, T, N, ts) : emo_set_library(ctx, colors, tile_px, T, N, ts);
    const NcclApi *api = nccl();
    if (!api) return EMO_ERR_NCCL;
    ncclComm_t comm = (ncclComm_t)ctx->comm;
    EMO_CK(cudaSetDevice(ctx->device));
    const bool is_root = ctx->rank == root;
    // header: the root's sizes, and whether the root accepted its own arguments (a rank that bailed out alone would leave
    // the others waiting in the broadcast)
    uint32_t hdr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int root_rc = EMO_OK;
    if (is_root) {
        root_rc = emo_check_library_args(colors, T, N, ts, tile_px);
        hdr[0] = T; hdr[1] = N; hdr[2] = ts; hdr[3] = tile_px != nullptr; hdr[4] = root_rc == EMO_OK;
    }
    const std::string root_msg = emo_last_error();
    if (!ctx->comm_hdr) EMO_CK(cudaMalloc(&ctx->comm_hdr, 64));
    if (is_root) EMO_CK(cudaMemcpyAsync(ctx->comm_hdr, hdr, sizeof hdr, cudaMemcpyHostToDevice, ctx->stream));
    EMO_NCCL(api, api->Broadcast(ctx->comm_hdr, ctx->comm_hdr, sizeof hdr, ncclUint8, root, comm, ctx->stream));
    EMO_CK(cudaMemcpyAsync(hdr, ctx->comm_hdr, sizeof hdr, cudaMemcpyDeviceToHost, ctx->stream));
    EMO_CK(cudaStreamSynchronize(ctx->stream));
    if (!hdr[4]) {
        if (is_root) emo_set_error("%s", root_msg.c_str());
        else emo_set_error("emo_comm_set_library: rank %d (the root) rejected its library arguments", root);
        return is_root ? root_rc : EMO_ERR_ARG;
    }
    T = hdr[0]; N = hdr[1]; ts = hdr[2];
    const bool has_px = hdr[3] != 0;
    ctx->T = 0;
    int rc = emo_library_common(ctx, T, N, ts, has_px);
    if (rc) { ctx->T = 0; return rc; }
    const size_t cb = (size_t)T * N * 3, pb = has_px ? (size_t)T * ts * ts * 3 : 0;
    uint8_t *c_buf, *p_buf = nullptr;
    if (is_root && dev_ptrs) {  // broadcast in place from the caller's device buffers
        c_buf = const_cast<uint8_t *>(colors);
        p_buf = const_cast<uint8_t *>(tile_px);
    } else {
        if ((rc = emo_ensure(ctx, &ctx->stage[0], &ctx->stage_cap[0], cb))) return rc;
        c_buf = (uint8_t *)ctx->stage[0];
        if (has_px) {
            if ((rc = emo_ensure(ctx, &ctx->stage[1], &ctx->stage_cap[1], pb))) return rc;
            p_buf = (uint8_t *)ctx->stage[1];
        }
        if (is_root) {
            EMO_CK(cudaMemcpyAsync(c_buf, colors, cb, cudaMemcpyHostToDevice, ctx->stream));
            if (has_px) EMO_CK(cudaMemcpyAsync(p_buf, tile_px, pb, cudaMemcpyHostToDevice, ctx->stream));
        }
    }
    EMO_NCCL(api, api->GroupStart());
    ncclResult_t r1 = api->Broadcast(c_buf, c_buf, cb, ncclUint8, root, comm, ctx->stream);
    ncclResult_t r2 = has_px ? api->Broadcast(p_buf, p_buf, pb, ncclUint8, root, comm, ctx->stream) : ncclSuccess;
    ncclResult_t r3 = api->GroupEnd();
    for (ncclResult_t r : {r1, r2, r3})
        if (r != ncclSuccess) {
            emo_set_error("ncclBroadcast of the library failed: %s", api->GetErrorString(r));
            return EMO_ERR_NCCL;
        }
    if ((rc = emo_launch_build_library(ctx, c_buf, p_buf))) return rc;
    if (!dev_ptrs) EMO_CK(cudaStreamSynchronize(ctx->stream));
    return EMO_OK;
}

int emo_comm_set_library(emo_ctx *ctx, const uint8_t *colors, const uint8_t *tile_px, uint32_t T, uint32_t N, uint32_t ts, int root) {
    return comm_set_library(ctx, colors, tile_px, T, N, ts, root, false);
}
int emo_comm_set_library_dev(emo_ctx *ctx, const uint8_t *colors, const uint8_t *tile_px, uint32_t T, uint32_t N, uint32_t ts,
                             int root) {
    return comm_set_library(ctx, colors, tile_px, T, N, ts, root, true);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------
// (b) one process, several GPUs
// ---------------------------------------------------------------------------------------
struct emo_group {
    std::vector<emo_ctx *> ctx;
    std::vector<ncclComm_t> comms;
};

// f(i) on one worker thread per member (member 0 on the calling thread); the first failure is reported with its GPU
template <typename F>
static int group_run(emo_group *g, F f) {
    const int n = (int)g->ctx.size();
    std::vector<int> rc(n, EMO_OK);
    std::vector<std::string> msg(n);
    auto work = [&](int i) {
        rc[i] = f(i);
        if (rc[i]) msg[i] = emo_last_error();  // thread-local: fetch it on the thread that failed
    };
    std::vector<std::thread> th;
    for (int i = 1; i < n; i++) th.emplace_back(work, i);
    work(0);
    for (std::thread &t : th) t.join();
    for (int i = 0; i < n; i++)
        if (rc[i]) {
            emo_set_error("GPU %d (member %d of %d): %s", g->ctx[i]->device, i, n, msg[i].c_str());
            return rc[i];
        }
    return EMO_OK;
}

extern "C" {

void emo_group_destroy(emo_group *g) {
    if (!g) return;
    for (emo_ctx *c : g->ctx) {
        if (!c) continue;
        cudaSetDevice(c->device);
        cudaDeviceSynchronize();
    }
    for (ncclComm_t c : g->comms)
        if (c && g_nccl.CommDestroy) g_nccl.CommDestroy(c);
    for (emo_ctx *c : g->ctx) {
        if (!c) continue;
        c->comm = nullptr;
        emo_destroy(c);
    }
    delete g;
}

int emo_group_create(const int *devices, int n, emo_group **out) {
    EMO_REQUIRE(out, EMO_ERR_ARG, "emo_group_create: out is NULL");
    *out = nullptr;
    EMO_REQUIRE(n >= 1 && n <= 64, EMO_ERR_ARG, "emo_group_create: %d devices", n);
    std::vector<int> devs(n);
    for (int i = 0; i < n; i++) {
        devs[i] = devices ? devices[i] : i;
        for (int j = 0; j < i; j++) EMO_REQUIRE(devs[j] != devs[i], EMO_ERR_ARG, "emo_group_create: device %d listed twice", devs[i]);
    }
    emo_group *g = new emo_group();
    g->ctx.assign(n, nullptr);
    const int rc = [&]() -> int {
        for (int i = 0; i < n; i++) {
            int r = emo_create(devs[i], &g->ctx[i]);
            if (r) return r;
        }
        if (n > 1) {
            const NcclApi *api = nccl();
            if (!api) return EMO_ERR_NCCL;
            g->comms.assign(n, nullptr);
            EMO_NCCL(api, api->CommInitAll(g->comms.data(), n, devs.data()));
            for (int i = 0; i < n; i++) {
                g->ctx[i]->comm = g->comms[i];
                g->ctx[i]->comm_owned = false;
                g->ctx[i]->rank = i;
                g->ctx[i]->world = n;
            }
        }
        return EMO_OK;
    }();
    if (rc) {
        const std::string msg = emo_last_error();
        emo_group_destroy(g);
        emo_set_error("%s", msg.c_str());
        return rc;
    }
    *out = g;
    return EMO_OK;
}

int emo_group_size(const emo_group *g) { return g ? (int)g->ctx.size() : 0; }

emo_ctx *emo_group_ctx(emo_group *g, int i) { return (g && i >= 0 && i < (int)g->ctx.size()) ? g->ctx[i] : nullptr; }

int emo_group_set_library(emo_group *g, const uint8_t *colors, const uint8_t *tile_px, uint32_t T, uint32_t N, uint32_t ts) {
    EMO_REQUIRE(g, EMO_ERR_ARG, "emo_group_set_library: group is NULL");
    const int n = (int)g->ctx.size();
    if (n == 1) return emo_set_library(g->ctx[0], colors, tile_px, T, N, ts);
    int rc = emo_check_library_args(colors, T, N, ts, tile_px);
    if (rc) return rc;
    const NcclApi *api = nccl();
    if (!api) return EMO_ERR_NCCL;
    const bool has_px = tile_px != nullptr;
    const size_t cb = (size_t)T * N * 3, pb = has_px ? (size_t)T * ts * ts * 3 : 0;
    for (int i = 0; i < n; i++) {
        emo_ctx *c = g->ctx[i];
        EMO_CK(cudaSetDevice(c->device));
        c->T = 0;
        if ((rc = emo_library_common(c, T, N, ts, has_px))) { c->T = 0; return rc; }
        if ((rc = emo_ensure(c, &c->stage[0], &c->stage_cap[0], cb))) return rc;
        if (has_px && (rc = emo_ensure(c, &c->stage[1], &c->stage_cap[1], pb))) return rc;
    }
    emo_ctx *c0 = g->ctx[0];
    EMO_CK(cudaSetDevice(c0->device));
    EMO_CK(cudaMemcpyAsync(c0->stage[0], colors, cb, cudaMemcpyHostToDevice, c0->stream));
    if (has_px) EMO_CK(cudaMemcpyAsync(c0->stage[1], tile_px, pb, cudaMemcpyHostToDevice, c0->stream));
    // one fused group: every member's colour and pixel broadcast (a single thread drives all communicators)
    EMO_NCCL(api, api->GroupStart());
    ncclResult_t bad = ncclSuccess;
    for (int i = 0; i < n && bad == ncclSuccess; i++) {
        emo_ctx *c = g->ctx[i];
        bad = api->Broadcast(c->stage[0], c->stage[0], cb, ncclUint8, 0, g->comms[i], c->stream);
        if (bad == ncclSuccess && has_px) bad = api->Broadcast(c->stage[1], c->stage[1], pb, ncclUint8, 0, g->comms[i], c->stream);
    }
    const ncclResult_t end = api->GroupEnd();
    if (bad == ncclSuccess) bad = end;
    if (bad != ncclSuccess) {
        emo_set_error("ncclBroadcast of the library failed: %s", api->GetErrorString(bad));
        return EMO_ERR_NCCL;
    }
    for (int i = 0; i < n; i++) {
        emo_ctx *c = g->ctx[i];
        EMO_CK(cudaSetDevice(c->device));
        if ((rc = emo_launch_build_library(c, (const uint8_t *)c->stage[0], has_px ? (const uint8_t *)c->stage[1] : nullptr))) return rc;
    }
    for (int i = 0; i < n; i++) {
        EMO_CK(cudaSetDevice(g->ctx[i]->device));
        EMO_CK(cudaStreamSynchronize(g->ctx[i]->stream));
    }
    return EMO_OK;
}

static int group_analyse(emo_group *g, const uint8_t *tiles, uint64_t T, uint32_t ts, uint32_t dim, uint8_t *out1, uint8_t *out4,
                         bool fused) {
    const int n = (int)g->ctx.size();
    const size_t tile_b = (size_t)ts * ts * 3, o1 = fused ? 3 : (size_t)dim * dim * 3;
    return group_run(g, [&](int i) -> int {
        uint64_t a, b;
        emo_stripe_bounds(T, n, i, &a, &b);
        if (a == b) return EMO_OK;
        emo_ctx *c = g->ctx[i];
        EMO_CK(cudaSetDevice(c->device));
        return emo_analyse_host_impl(c, tiles + a * tile_b, b - a, ts, dim, out1 + a * o1, fused ? out4 + a * 12 : nullptr, fused);
    });
}

int emo_group_analyse(emo_group *g, const uint8_t *tiles, uint64_t T, uint32_t ts, uint32_t dim, uint8_t *out) {
    EMO_REQUIRE(g, EMO_ERR_ARG, "emo_group_analyse: group is NULL");
    int rc = emo_check_analyse_args(tiles, T, ts, dim, out);
    if (rc || T == 0) return rc;
    return group_analyse(g, tiles, T, ts, dim, out, nullptr, false);
}

int emo_group_analyse_fused(emo_group *g, const uint8_t *tiles, uint64_t T, uint32_t ts, uint8_t *out1, uint8_t *out4) {
    EMO_REQUIRE(g, EMO_ERR_ARG, "emo_group_analyse_fused: group is NULL");
    int rc = emo_check_analyse_args(tiles, T, ts, 2, out1);
    if (rc) return rc;
    EMO_REQUIRE(T == 0 || out4, EMO_ERR_ARG, "analyse_fused: out4 is NULL");
    EMO_REQUIRE(ts % 2 == 0, EMO_ERR_ARG, "Invalid tile size: Tile size must be divisible by 2");  // main.rs:612-615
    if (T == 0) return EMO_OK;
    return group_analyse(g, tiles, T, ts, 2, out1, out4, true);
}

int emo_group_mosaic(emo_group *g, const uint8_t *src, uint32_t W, uint32_t H, uint32_t oc, uint8_t tint_alpha, int32_t *item,
                     uint32_t *dist, uint8_t *out) {
    EMO_REQUIRE(g, EMO_ERR_ARG, "emo_group_mosaic: group is NULL");
    const int n = (int)g->ctx.size();
    emo_ctx *c0 = g->ctx[0];
    if (n == 1) return emo_mosaic_host_impl(c0, src, W, H, oc, tint_alpha, item, dist, out, 0);
    EMO_REQUIRE(c0->T > 0, EMO_ERR_STATE, "mosaic: no library set (call emo_group_set_library first)");
    EMO_REQUIRE(src && out, EMO_ERR_ARG, "mosaic: src/out is NULL");
    const uint32_t dim = c0->dim, tsz = c0->ts;
    // main.rs:603-611
    EMO_REQUIRE(W > 0 && H > 0 && W % dim == 0 && H % dim == 0, EMO_ERR_ARG,
                "Invalid source dimensions (%ux%u): Dimensions must be divisible by %u", W, H, dim);
    const uint32_t bw = W / dim, bh = H / dim;
    const uint64_t Q = (uint64_t)bw * bh;
    const size_t row_out = (size_t)bw * tsz * tsz * oc, row_src = (size_t)dim * W * 3;
    return group_run(g, [&](int i) -> int {
        uint64_t a, b;
        emo_stripe_bounds(bh, n, i, &a, &b);
        if (a == b) return EMO_OK;  // more GPUs than block rows
        return emo_mosaic_host_impl(g->ctx[i], src + a * row_src, W, (uint32_t)(b - a) * dim, oc, tint_alpha,
                                    item ? item + a * bw : nullptr, dist ? dist + a * bw : nullptr, out + a * row_out, Q);
    });
}

}  // extern "C"
