// stats.cu — the reductions of RenderStats over the maps the match already left on the device.
//
// Reference: render_nto1 pushes one entry per block into a Mutex<RenderStats> (src/mosaic/rendering.rs:211-214 ->
// stats.rs:56-64) and summarise() walks that map (stats.rs:87-139): number of tiles placed, total / average distance, how
// often each tile was used (top 10), the worst matches; render() needs the maximum distance (stats.rs:169-175).  The counts
// and sums are one pass over item / dist; only the two top-10 lists need an order and stay with the host layers.
#include "common.cuh"

// sums[0] = blocks with a tile, sums[1] = sum of their distances, sums[2] = their maximum distance;
// usage[t] += 1 for every block that placed tile t + 1 in either orientation (tile.idx is unsigned, stats.rs:57-62)
__global__ void __launch_bounds__(256) stats_kernel(const int32_t *__restrict__ item, const uint32_t *__restrict__ dist, uint64_t Q, uint32_t T,
                                                    unsigned long long *__restrict__ sums, uint32_t *__restrict__ usage, int *__restrict__ err) {
    unsigned long long placed = 0, total = 0;
    uint32_t mx = 0;
    for (uint64_t q = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; q < Q; q += (uint64_t)gridDim.x * blockDim.x) {
        const int32_t it = item[q];
        if (it == 0) continue;  // a no-repeat block that ran out of tiles has no entry (rendering.rs:347-365)
        const uint32_t t = (uint32_t)(it < 0 ? -it : it) - 1;
        if (t >= T) { atomicOr(err, 1); continue; }
        const uint32_t d = dist[q];
        placed++;
        total += d;
        mx = max(mx, d);
        if (usage) atomicAdd(&usage[t], 1u);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        placed += __shfl_xor_sync(0xFFFFFFFFu, placed, o);
        total += __shfl_xor_sync(0xFFFFFFFFu, total, o);
        mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
    }
    if ((threadIdx.x & 31) == 0 && placed) {
        atomicAdd(&sums[0], placed);
        atomicAdd(&sums[1], total);
        atomicMax(&sums[2], (unsigned long long)mx);
    }
}

static int launch_stats(emo_ctx *ctx, const int32_t *item, const uint32_t *dist, uint64_t Q, uint32_t T, unsigned long long *sums,
                        uint32_t *usage) {
    EMO_CK(cudaMemsetAsync(sums, 0, 3 * sizeof(unsigned long long), ctx->stream));
    if (usage) EMO_CK(cudaMemsetAsync(usage, 0, (size_t)T * 4, ctx->stream));
    if (Q == 0) return EMO_OK;
    const uint64_t want = (Q + 1023) / 1024;
    const uint32_t blocks = (uint32_t)(want < (uint64_t)ctx->sm_count * 8 ? want : (uint64_t)ctx->sm_count * 8);
    stats_kernel<<<blocks, 256, 0, ctx->stream>>>(item, dist, Q, T, sums, usage, ctx->err_flag);
    EMO_LAUNCH_CHECK(ctx);
    return EMO_OK;
}

extern "C" {

int emo_stats_dev(emo_ctx *ctx, const int32_t *item_dev, const uint32_t *dist_dev, uint64_t Q, uint32_t T, uint64_t *sums_dev,
                  uint32_t *usage_dev) {
    EMO_REQUIRE(ctx && sums_dev, EMO_ERR_ARG, "emo_stats_dev: NULL argument");
    EMO_REQUIRE(Q == 0 || (item_dev && dist_dev), EMO_ERR_ARG, "emo_stats_dev: NULL map");
    EMO_REQUIRE((uintptr_t)sums_dev % 8 == 0, EMO_ERR_ARG, "emo_stats_dev: sums must be 8-byte aligned");
    EMO_CK(cudaSetDevice(ctx->device));
    return launch_stats(ctx, item_dev, dist_dev, Q, T, (unsigned long long *)sums_dev, usage_dev);
}

int emo_stats(emo_ctx *ctx, const int32_t *item, const uint32_t *dist, uint64_t Q, uint32_t T, uint64_t sums[3], uint32_t *usage) {
    EMO_REQUIRE(ctx && sums, EMO_ERR_ARG, "emo_stats: NULL argument");
    EMO_REQUIRE(Q == 0 || (item && dist), EMO_ERR_ARG, "emo_stats: NULL map");
    EMO_CK(cudaSetDevice(ctx->device));
    int rc;
    const size_t ub = usage ? (size_t)T * 4 : 0;
    if ((rc = emo_ensure(ctx, &ctx->stage[3], &ctx->stage_cap[3], Q * 4))) return rc;
    if ((rc = emo_ensure(ctx, &ctx->stage[4], &ctx->stage_cap[4], Q * 4))) return rc;
    if ((rc = emo_ensure(ctx, &ctx->stage[0], &ctx->stage_cap[0], 32 + ub))) return rc;
    if (Q) {
        EMO_CK(cudaMemcpyAsync(ctx->stage[3], item, Q * 4, cudaMemcpyHostToDevice, ctx->stream));
        EMO_CK(cudaMemcpyAsync(ctx->stage[4], dist, Q * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    unsigned long long *d_sums = (unsigned long long *)ctx->stage[0];
    uint32_t *d_usage = usage ? (uint32_t *)((uint8_t *)ctx->stage[0] + 32) : nullptr;
    if ((rc = launch_stats(ctx, (const int32_t *)ctx->stage[3], (const uint32_t *)ctx->stage[4], Q, T, d_sums, d_usage))) return rc;
    EMO_CK(cudaMemcpyAsync(sums, d_sums, 24, cudaMemcpyDeviceToHost, ctx->stream));
    if (usage) EMO_CK(cudaMemcpyAsync(usage, d_usage, ub, cudaMemcpyDeviceToHost, ctx->stream));
    EMO_CK(cudaStreamSynchronize(ctx->stream));
    return emo_check_device_flag(ctx);
}

}  // extern "C"
