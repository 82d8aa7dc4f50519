// hostcopy.cu — what the host side of the box can absorb: pinned-memory copy ceiling with n GPUs copying at once.
// Measurement code (tools/libemosaic_probe.so), used by bench.py to put the end-to-end number next to the rate that the
// PCIe links + host memory deliver when every GPU drains its output stripe at the same time.
//
//   int emo_probe_host_copy(const int *devices, int n, size_t bytes, size_t chunk, int dir, int reps,
//                           double *aggregate_gbs, double *per_device_gbs /* [n] */)
//   Every device gets its own worker thread, its own device buffer and its own pinned host buffer (cudaHostAlloc from the
//   worker thread, touched before use), and copies `bytes` `reps` times in pieces of `chunk` bytes — one cudaMemcpyAsync
//   per piece, like emo_mosaic's drain.  dir 0 = device -> host, 1 = host -> device.  All workers start together;
//   aggregate = n * bytes * reps / (wall time from the common start to the last worker's finish); per-device figures are
//   CUDA-event times of each worker's own copies.  Returns 0 or a negative cudaError.
//   emo_probe_host_copy_at(..., start_unix_ns): the same, but the timed copies start at that wall-clock instant
//   (CLOCK_REALTIME), so that several PROCESSES on one box (one rank per GPU) copy at the same time; returns 1 when the
//   set-up (allocation, warm-up pass) finished after the instant — the figure would not be a concurrent one.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

#include <time.h>
static int64_t unix_ns() {
    timespec ts;
    clock_gettime(CLOCK_REALTIME, &ts);
    return (int64_t)ts.tv_sec * 1000000000ll + ts.tv_nsec;
}

extern "C" int emo_probe_host_copy_at(const int *devices, int n, size_t bytes, size_t chunk, int dir, int reps, double *aggregate_gbs,
                                      double *per_device_gbs, int64_t start_unix_ns) {
    if (n < 1 || n > 64 || bytes == 0 || reps < 1) return -(int)cudaErrorInvalidValue;
    if (chunk == 0 || chunk > bytes) chunk = bytes;
    std::vector<int> rc(n, 0);
    std::vector<double> ms(n, 0.0);
    std::atomic<int> ready{0}, go{0};
    std::vector<std::chrono::steady_clock::time_point> done(n);
    auto work = [&](int i) {
        const int dev = devices ? devices[i] : i;
        void *d = nullptr, *h = nullptr;
        cudaStream_t st = nullptr;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        auto fail = [&](cudaError_t e) { rc[i] = -(int)e; };
        cudaError_t e;
        if ((e = cudaSetDevice(dev)) != cudaSuccess || (e = cudaMalloc(&d, bytes)) != cudaSuccess ||
            (e = cudaHostAlloc(&h, bytes, cudaHostAllocDefault)) != cudaSuccess ||
            (e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking)) != cudaSuccess || (e = cudaEventCreate(&e0)) != cudaSuccess ||
            (e = cudaEventCreate(&e1)) != cudaSuccess)
            fail(e);
        if (!rc[i]) {
            memset(h, 1, bytes);
            cudaMemsetAsync(d, 2, bytes, st);
            // warm-up: one full pass
            if (dir == 0) cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, st);
            else cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st);
            if ((e = cudaStreamSynchronize(st)) != cudaSuccess) fail(e);
        }
        ready.fetch_add(1);
        while (go.load() == 0) std::this_thread::yield();
        if (!rc[i]) {
            cudaEventRecord(e0, st);
            for (int r = 0; r < reps; r++)
                for (size_t o = 0; o < bytes; o += chunk) {
                    const size_t m = bytes - o < chunk ? bytes - o : chunk;
                    if (dir == 0) cudaMemcpyAsync((char *)h + o, (char *)d + o, m, cudaMemcpyDeviceToHost, st);
                    else cudaMemcpyAsync((char *)d + o, (char *)h + o, m, cudaMemcpyHostToDevice, st);
                }
            cudaEventRecord(e1, st);
            if ((e = cudaEventSynchronize(e1)) != cudaSuccess) fail(e);
            float t = 0;
            cudaEventElapsedTime(&t, e0, e1);
            ms[i] = t;
        }
        done[i] = std::chrono::steady_clock::now();
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        if (st) cudaStreamDestroy(st);
        if (h) cudaFreeHost(h);
        if (d) cudaFree(d);
    };
    std::vector<std::thread> th;
    for (int i = 0; i < n; i++) th.emplace_back(work, i);
    while (ready.load() < n) std::this_thread::yield();
    bool late = false;
    if (start_unix_ns) {
        late = unix_ns() > start_unix_ns;
        while (unix_ns() < start_unix_ns) {
        }
    }
    const auto t0 = std::chrono::steady_clock::now();
    go.store(1);
    for (auto &t : th) t.join();
    for (int i = 0; i < n; i++)
        if (rc[i]) return rc[i];
    double wall = 0;
    for (int i = 0; i < n; i++) {
        const double s = std::chrono::duration<double>(done[i] - t0).count();
        if (s > wall) wall = s;
        if (per_device_gbs) per_device_gbs[i] = (double)bytes * reps / (ms[i] * 1e-3) / 1e9;
    }
    if (aggregate_gbs) *aggregate_gbs = (double)n * bytes * reps / wall / 1e9;
    return late ? 1 : 0;
}

extern "C" int emo_probe_host_copy(const int *devices, int n, size_t bytes, size_t chunk, int dir, int reps, double *aggregate_gbs,
                                   double *per_device_gbs) {
    return emo_probe_host_copy_at(devices, n, bytes, chunk, dir, reps, aggregate_gbs, per_device_gbs, 0);
}
