// norepeat.cu — emo_no_repeat: the assignment of render_nto1_no_repeat (src/mosaic/rendering.rs:262-392) from one call.
//
// The reference scores every block against the whole search set (nearest_n(100000), :307-321), keeps the blocks sorted by the
// distance of their best remaining candidate (:323-326), pops the nearest, places its tile if nobody used it (either
// orientation, :357-358) and otherwise moves the block to its next candidate (:380-391), re-querying the pruned tree when a
// list runs out (:384-386).  That loop is a merge of the per-block sorted candidate lists in globally increasing
// (distance, block number) order; it is sequential, but it only ever looks at the head of each list.
//
// Here the lists come from the GPU in pages (topk_kernel, topk.cu): a deep first page for every block in one launch, sized
// so that running dry is rare (up to 1024 candidates per block within a 512 MB host budget), and the merge runs on the host
// over 64-bit heap keys.  A block that runs dry is parked; its next candidate cannot be nearer than the one it just lost, so
// the merge continues with every block below that bound and refills all parked blocks with ONE launch (the nearest
// candidates among the tiles still free: what the reference's pruned tree returns) when the bound is reached.
#include <string.h>

#include <algorithm>
#include <queue>

#include "common.cuh"

extern "C" int emo_no_repeat(emo_ctx *ctx, const uint8_t *src, uint32_t W, uint32_t H, uint32_t page, int32_t *item, uint32_t *dist,
                             uint64_t *counters) {
    EMO_REQUIRE(ctx, EMO_ERR_ARG, "emo_no_repeat: ctx is NULL");
    EMO_REQUIRE(ctx->T > 0, EMO_ERR_STATE, "no_repeat: no library set (call emo_set_library first)");
    EMO_REQUIRE(src && item && dist, EMO_ERR_ARG, "no_repeat: NULL buffer");
    EMO_REQUIRE(W > 0 && H > 0, EMO_ERR_ARG, "no_repeat: empty source %ux%u", W, H);
    const uint32_t dim = ctx->dim, T = ctx->T;
    EMO_REQUIRE(W % dim == 0 && H % dim == 0, EMO_ERR_ARG, "Invalid source dimensions (%ux%u): Dimensions must be divisible by %u", W, H,
                dim);  // main.rs:603-611
    if (ctx->wide) {
        emo_set_error("no_repeat: ranked candidate lists exist for --mode 1..4 (N = 1, 4, 9, 16), not N=%u", ctx->N);
        return EMO_ERR_UNSUPPORTED;
    }
    EMO_REQUIRE(page <= 1024, EMO_ERR_ARG, "no_repeat: page=%u outside [0,1024]", page);
    const uint32_t bw = W / dim, bh = H / dim;
    const uint64_t Q64 = (uint64_t)bw * bh;
    // rendering.rs:292-298
    EMO_REQUIRE(Q64 <= 2ull * T, EMO_ERR_ARG, "Insufficient tiles for no-repeat mode: need %llu tiles but only have %llu available",
                (unsigned long long)Q64, 2ull * T);
    const uint32_t Q = (uint32_t)Q64;
    EMO_CK(cudaSetDevice(ctx->device));
    // first page: as deep as a 512 MB host budget allows (8 bytes per entry), at most the whole list
    const uint32_t L = ctx->L;
    uint32_t k0 = page ? page : (uint32_t)std::min<uint64_t>(1024, std::max<uint64_t>(16, (64ull << 20) / Q));
    if (k0 > L) k0 = std::max<uint32_t>(L, 1);
    std::vector<int32_t> p_item((size_t)Q * k0);
    std::vector<uint32_t> p_dist((size_t)Q * k0);
    int rc = emo_topk(ctx, src, W, H, 0, k0, nullptr, p_item.data(), p_dist.data());
    if (rc) return rc;

    // per block: where its current page lives (the first pages in p_*, refills in r_*), its length and read position
    struct Blk {
        const int32_t *it;
        const uint32_t *ds;
        uint32_t len, ptr;
    };
    std::vector<Blk> blk(Q);
    // heap key = distance << 32 | n with the reference's block number n = bx * vtiles + by (:300-301): unique per block, so the
    // key order IS the canonical (distance, n) order and the block is recovered from n
    std::priority_queue<uint64_t, std::vector<uint64_t>, std::greater<uint64_t>> heap;
    {
        std::vector<uint64_t> init;
        init.reserve(Q);
        for (uint32_t q = 0; q < Q; q++) {
            blk[q] = Blk{p_item.data() + (size_t)q * k0, p_dist.data() + (size_t)q * k0, k0, 0};
            const uint32_t by = q / bw, bx = q % bw;
            if (blk[q].it[0] != 0) init.push_back((uint64_t)blk[q].ds[0] << 32 | (uint64_t)(bx * bh + by));
        }
        heap = std::priority_queue<uint64_t, std::vector<uint64_t>, std::greater<uint64_t>>(std::greater<uint64_t>(), std::move(init));
    }
    memset(item, 0, (size_t)Q * 4);
    memset(dist, 0, (size_t)Q * 4);
    std::vector<uint8_t> retired(T, 0);  // tiles placed so far: what the reference removes from the tree (:366-380)
    uint32_t n_retired = 0;
    std::vector<std::pair<uint64_t, uint32_t>> parked;  // (bound key, block)
    uint64_t parked_bound = ~0ull;
    std::vector<std::vector<int32_t>> r_item;  // refill pages stay alive until the end (blocks point into them)
    std::vector<std::vector<uint32_t>> r_dist;
    std::vector<uint8_t> strip;
    uint64_t n_refills = 0, n_refilled = 0, n_pops = 0;

    auto refill = [&]() -> int {
        // one launch for every parked block: a strip image with one block per parked entry, lists filtered by `retired`
        const uint32_t m = (uint32_t)parked.size();
        uint32_t k = 0;
        for (auto &p : parked) k = std::max(k, std::min<uint32_t>(1024, std::max<uint32_t>(64, 2 * blk[p.second].len)));
        k = std::min<uint32_t>(k, std::max<uint32_t>(L, 1));
        strip.resize((size_t)m * dim * dim * 3);
        for (uint32_t i = 0; i < m; i++) {
            const uint32_t q = parked[i].second, by = q / bw, bx = q % bw;
            for (uint32_t row = 0; row < dim; row++)
                memcpy(&strip[((size_t)row * m * dim + (size_t)i * dim) * 3], src + ((size_t)(by * dim + row) * W + (size_t)bx * dim) * 3,
                       (size_t)dim * 3);
        }
        r_item.emplace_back((size_t)m * k);
        r_dist.emplace_back((size_t)m * k);
        int rc2 = emo_topk(ctx, strip.data(), m * dim, dim, 0, k, retired.data(), r_item.back().data(), r_dist.back().data());
        if (rc2) return rc2;
        n_refills++;
        n_refilled += m;
        for (uint32_t i = 0; i < m; i++) {
            const uint32_t q = parked[i].second, by = q / bw, bx = q % bw;
            blk[q] = Blk{r_item.back().data() + (size_t)i * k, r_dist.back().data() + (size_t)i * k, k, 0};
            if (blk[q].it[0] != 0) heap.push((uint64_t)blk[q].ds[0] << 32 | (uint64_t)(bx * bh + by));  // else: nothing left, stays black
        }
        parked.clear();
        parked_bound = ~0ull;
        return EMO_OK;
    };

    while (!heap.empty() || !parked.empty()) {
        if (!parked.empty() && (heap.empty() || heap.top() >= parked_bound)) {
            if (n_retired >= T) break;  // out of tiles: the parked blocks stay black (:347-351)
            if ((rc = refill())) return rc;
            continue;
        }
        const uint64_t key = heap.top();
        heap.pop();
        n_pops++;
        const uint32_t n = (uint32_t)key, d = (uint32_t)(key >> 32);
        const uint32_t bx = n / bh, by = n % bh, q = by * bw + bx;
        Blk &b = blk[q];
        const int32_t it = b.it[b.ptr];
        const uint32_t a = (uint32_t)(it < 0 ? -it : it) - 1;
        if (!retired[a]) {
            retired[a] = 1;
            n_retired++;
            item[q] = it;
            dist[q] = d;
            continue;
        }
        if (++b.ptr >= b.len) {
            if (b.len >= L) continue;  // the page was the whole list: nothing left for this block
            // dry: the next candidate is at least as far as the one just lost, and the block keeps its number
            parked.emplace_back(key, q);
            parked_bound = std::min(parked_bound, key);
            continue;
        }
        if (b.it[b.ptr] == 0) continue;  // end of the list
        heap.push((uint64_t)b.ds[b.ptr] << 32 | n);
    }
    if (counters) {
        counters[0] = n_refills;
        counters[1] = n_refilled;
        counters[2] = n_pops;
        counters[3] = n_retired;
    }
    return EMO_OK;
}
