// cs_stage.cu — pageable host memory <-> device copies at PCIe speed.
//
// The host entry points of the C ABI take R's own vectors (pageable memory).  cudaMemcpy on
// pageable memory stages through one small driver buffer on one thread: ~5 GB/s measured on the
// B200 box, a fifth of the link.  Here a copy is cut into pieces; a few worker threads each own two
// pinned slots and a stream (three quarters of the host's cores, at most 12), so that one thread's memcpy between pageable memory and its slot runs
// while the DMA of its other slot (and of every other thread's slots) is in flight.  The pinned
// slots are allocated once per process and reused (page-locking memory costs more than the copy).
#include "cs_common.cuh"

#include <algorithm>
#include <atomic>
#include <mutex>
#include <string.h>
#include <thread>
#include <vector>

namespace {

constexpr size_t kPiece = 4u << 20;   // bytes per pinned slot
constexpr int kMaxWorkers = 12;
constexpr size_t kSmall = 8u << 20;   // below this a plain cudaMemcpy is as fast

struct Pool {
    std::mutex mu;                    // one staged copy at a time per process
    int device = -1;
    void *slot[kMaxWorkers][2] = {};
    cudaStream_t st[kMaxWorkers] = {};
    cudaEvent_t ev[kMaxWorkers][2] = {};
    bool ready = false;
    cudaError_t init(int dev)
    {
        if (ready && dev == device) return cudaSuccess;
        release();
        cudaError_t e;
        for (int w = 0; w < kMaxWorkers; w++) {
            if ((e = cudaStreamCreateWithFlags(&st[w], cudaStreamNonBlocking)) != cudaSuccess) return e;
            for (int b = 0; b < 2; b++) {
                if ((e = cudaHostAlloc(&slot[w][b], kPiece, cudaHostAllocDefault)) != cudaSuccess) return e;
                if ((e = cudaEventCreateWithFlags(&ev[w][b], cudaEventDisableTiming)) != cudaSuccess) return e;
            }
        }
        device = dev;
        ready = true;
        return cudaSuccess;
    }
    void release()
    {
        for (int w = 0; w < kMaxWorkers; w++) {
            for (int b = 0; b < 2; b++) {
                if (slot[w][b]) cudaFreeHost(slot[w][b]);
                if (ev[w][b]) cudaEventDestroy(ev[w][b]);
                slot[w][b] = nullptr;
                ev[w][b] = nullptr;
            }
            if (st[w]) cudaStreamDestroy(st[w]);
            st[w] = nullptr;
        }
        ready = false;
    }
};

Pool g_pool;

int workers_for(size_t bytes)
{
    unsigned hw = std::thread::hardware_concurrency();
    int w = hw ? (int)std::min<unsigned>(kMaxWorkers, std::max(1u, hw * 3 / 4)) : 4; // (measured: 4 / 8 / 12 / 16 threads = 20 / 25 / 28 / 28 GB/s on a 16-core host)
    if (const char *e = getenv("CS_STAGE_THREADS")) w = std::max(1, std::min(kMaxWorkers, atoi(e)));
    const size_t pieces = (bytes + kPiece - 1) / kPiece;
    return (int)std::min<size_t>(w, pieces);
}

// pieces w, w + W, w + 2W, ... of the copy, double-buffered over this worker's two slots
void worker(Pool *P, int dev, int w, int W, char *dev_ptr, char *host_ptr, size_t bytes, bool to_device,
            std::atomic<int> *err)
{
    if (cudaSetDevice(dev) != cudaSuccess) { err->store((int)cudaGetLastError()); return; }
    const size_t pieces = (bytes + kPiece - 1) / kPiece;
    auto span = [&](size_t i, size_t &off, size_t &len) {
        off = i * kPiece;
        len = std::min(kPiece, bytes - off);
    };
    cudaError_t e = cudaSuccess;
    if (to_device) {
        int b = 0;
        for (size_t i = w; i < pieces && e == cudaSuccess; i += W, b ^= 1) {
            size_t off, len;
            span(i, off, len);
            if ((e = cudaEventSynchronize(P->ev[w][b])) != cudaSuccess) break; // the slot's previous DMA is done
            memcpy(P->slot[w][b], host_ptr + off, len);
            if ((e = cudaMemcpyAsync(dev_ptr + off, P->slot[w][b], len, cudaMemcpyHostToDevice, P->st[w])) != cudaSuccess) break;
            e = cudaEventRecord(P->ev[w][b], P->st[w]);
        }
    } else {
        size_t prev = (size_t)-1;
        int b = 0;
        for (size_t i = w;; i += W, b ^= 1) {
            if (i < pieces) {
                size_t off, len;
                span(i, off, len);
                if ((e = cudaMemcpyAsync(P->slot[w][b], dev_ptr + off, len, cudaMemcpyDeviceToHost, P->st[w])) != cudaSuccess) break;
                if ((e = cudaEventRecord(P->ev[w][b], P->st[w])) != cudaSuccess) break;
            }
            if (prev != (size_t)-1) {
                size_t off, len;
                span(prev, off, len);
                if ((e = cudaEventSynchronize(P->ev[w][b ^ 1])) != cudaSuccess) break;
                memcpy(host_ptr + off, P->slot[w][b ^ 1], len);
            }
            if (i >= pieces) break;
            prev = i;
        }
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(P->st[w]);
    if (e != cudaSuccess) err->store((int)e);
}

int staged(void *dev_ptr, void *host_ptr, size_t bytes, bool to_device)
{
    if (bytes == 0) return CS_OK;
    if (bytes < kSmall) {
        CS_CUDA(cudaMemcpy(to_device ? dev_ptr : host_ptr, to_device ? host_ptr : dev_ptr, bytes,
                           to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost));
        // cudaMemcpy from pageable memory returns once the data is in the driver's staging buffer; the
        // DMA runs on the legacy stream, which the callers' non-blocking streams do not wait for
        if (to_device) CS_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
        return CS_OK;
    }
    int dev = 0;
    CS_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_pool.mu);
    CS_CUDA(g_pool.init(dev));
    const int W = workers_for(bytes);
    std::atomic<int> err{0};
    std::vector<std::thread> th;
    for (int w = 1; w < W; w++)
        th.emplace_back(worker, &g_pool, dev, w, W, (char *)dev_ptr, (char *)host_ptr, bytes, to_device, &err);
    worker(&g_pool, dev, 0, W, (char *)dev_ptr, (char *)host_ptr, bytes, to_device, &err);
    for (auto &t : th) t.join();
    if (err.load()) {
        cs_set_error("staged %s copy of %zu bytes failed: %s", to_device ? "host-to-device" : "device-to-host", bytes,
                     cudaGetErrorString((cudaError_t)err.load()));
        return CS_ERR_CUDA;
    }
    return CS_OK;
}

} // namespace

// Both return after the copy has completed.  The device side must not be in use by work that
// is still running (callers synchronise their stream first).
int cs_h2d_staged(void *d_dst, const void *h_src, size_t bytes) { return staged(d_dst, const_cast<void *>(h_src), bytes, true); }
int cs_d2h_staged(void *h_dst, const void *d_src, size_t bytes) { return staged(const_cast<void *>(d_src), h_dst, bytes, false); }
