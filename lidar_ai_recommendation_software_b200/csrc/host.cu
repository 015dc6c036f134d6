// Host side of the copies: NUMA placement, page-locked staging, parallel memcpy (include/lidar_b200.h).
//
// The drop-in surface takes caller-owned pageable numpy arrays and returns arrays the caller owns
// (SURVEY.md §8b "Ownership"; the reference keeps its outputs alive in st.session_state, app.py:84,204,247),
// so every frame crosses host memory twice on each side of PCIe: pageable -> page-locked -> device and back.
// One core copies ~10 GB/s; PCIe 5 x16 moves ~55 GB/s.  The staging copies therefore run on a small pool of
// worker threads, pinned (like the page-locked buffers) to the NUMA node the GPU hangs off — on the 8-GPU box
// every rank otherwise inherits the whole machine's affinity mask and its staging memory lands on node 0.
#include <dirent.h>
#include <pthread.h>
#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

#if defined(__x86_64__)
#include <emmintrin.h>
#endif

namespace lidar {

// ---- the copy itself -------------------------------------------------------------------------------
// Staging a frame writes 16 MB that the CPU never reads again (the DMA engine does).  Ordinary stores first pull every
// destination line into the cache (read-for-ownership): 48 MB of memory traffic for a 16 MB copy, while both DMA
// directions are using the same memory controllers.  Non-temporal stores write whole lines without reading them.
static std::atomic<int> g_copy_nt{1};

static void copy_bytes(char* d, const char* s, size_t n) {
#if defined(__x86_64__)
    if (g_copy_nt.load(std::memory_order_relaxed) && n >= (size_t)(256 << 10)) {
        size_t head = (64 - (reinterpret_cast<uintptr_t>(d) & 63)) & 63;
        memcpy(d, s, head);
        d += head; s += head; n -= head;
        const size_t lines = n / 64;
        for (size_t i = 0; i < lines; ++i, s += 64, d += 64) {
            const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s));
            const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 16));
            const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 32));
            const __m128i e = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 48));
            _mm_stream_si128(reinterpret_cast<__m128i*>(d), a);
            _mm_stream_si128(reinterpret_cast<__m128i*>(d + 16), b);
            _mm_stream_si128(reinterpret_cast<__m128i*>(d + 32), c);
            _mm_stream_si128(reinterpret_cast<__m128i*>(d + 48), e);
        }
        _mm_sfence();
        memcpy(d, s, n - lines * 64);
        return;
    }
#endif
    memcpy(d, s, n);
}

// ---- worker pool ---------------------------------------------------------------------------------
// A job is a list of (dst, src, bytes) segments treated as ONE byte range and cut into equal slices, one per worker
// plus one for the caller.  Workers spin on the job counter for a short while after each job before they go to sleep
// on the condition variable: the copies of a frame arrive in bursts (stage in, a DMA later copy out), and waking seven
// sleeping threads through the futex costs more than a 4 MB copy takes (measured on the 16-vCPU host: ~0.25 ms per
// wake-up against 0.22 ms for 28 MB at 126 GB/s).
struct CopySeg {
    char* dst;
    const char* src;
    size_t bytes;
};

class CopyPool {
  public:
    ~CopyPool() { stop(); }
    void resize(int n) {
        std::lock_guard<std::mutex> g(api_);
        if (n == (int)workers_.size()) return;
        stop();
        quit_.store(false);
        gen_.store(0);
        for (int i = 0; i < n; ++i) workers_.emplace_back([this, i] { run(i); });
    }
    int size() {
        std::lock_guard<std::mutex> g(api_);
        return (int)workers_.size();
    }
    // wake the workers without giving them work: they spin for the next job instead of sleeping through a DMA
    void wake() {
        std::lock_guard<std::mutex> g(api_);
        if (workers_.empty()) return;
        while (pending_.load(std::memory_order_acquire) != 0) cpu_relax();      // every generation is acknowledged before the next
        nseg_ = 0;
        total_ = 0;
        per_ = 64;
        pending_.store((int)workers_.size(), std::memory_order_relaxed);
        {
            std::lock_guard<std::mutex> l(m_);
            gen_.fetch_add(1, std::memory_order_release);
        }
        if (sleepers_.load(std::memory_order_acquire) > 0) cv_.notify_all();
        // no wait: the next job's pending_ store must not race with late decrements, so copy() waits for them first
    }
    void set_spin_us(int us) { spin_us_.store(us); }
    void copy(const CopySeg* segs, int nseg) {
        std::lock_guard<std::mutex> g(api_);
        while (pending_.load(std::memory_order_acquire) != 0) cpu_relax();      // a wake() still being acknowledged
        size_t total = 0;
        for (int k = 0; k < nseg; ++k) total += segs[k].bytes;
        if (workers_.empty() || total < (size_t)(1 << 20)) {
            for (int k = 0; k < nseg; ++k)
                if (segs[k].bytes) copy_bytes(segs[k].dst, segs[k].src, segs[k].bytes);
            return;
        }
        const size_t parts = workers_.size() + 1;
        segs_ = segs;
        nseg_ = nseg;
        total_ = total;
        per_ = ((total + parts - 1) / parts + 63) & ~(size_t)63;
        pending_.store((int)workers_.size(), std::memory_order_relaxed);
        {
            std::lock_guard<std::mutex> l(m_);          // a worker about to sleep re-checks gen_ under this lock
            gen_.fetch_add(1, std::memory_order_release);
        }
        if (sleepers_.load(std::memory_order_acquire) > 0) cv_.notify_all();
        slice((int)workers_.size());                    // the caller takes the last slice
        while (pending_.load(std::memory_order_acquire) != 0) cpu_relax();
    }

  private:
    static void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#else
        std::this_thread::yield();
#endif
    }
    // bytes [a, a + len) of the concatenated segments
    void slice(int i) {
        size_t a = (size_t)i * per_;
        if (a >= total_) return;
        size_t len = a + per_ <= total_ ? per_ : total_ - a;
        for (int k = 0; k < nseg_ && len; ++k) {
            const size_t sb = segs_[k].bytes;
            if (a >= sb) { a -= sb; continue; }
            const size_t take = sb - a < len ? sb - a : len;
            copy_bytes(segs_[k].dst + a, segs_[k].src + a, take);
            len -= take;
            a = 0;
        }
    }
    void run(int i) {
        unsigned long long seen = 0;
        for (;;) {
            // spin first (up to ~1 ms: longer than the DMA that usually separates a wake-up call from the copy), then sleep
            bool got = false;
            const auto t0 = std::chrono::steady_clock::now();
            for (int spin = 0;; ++spin) {
                if (quit_.load(std::memory_order_relaxed)) return;
                if (gen_.load(std::memory_order_acquire) != seen) { got = true; break; }
                cpu_relax();
                if ((spin & 255) == 255 &&
                    std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(spin_us_.load(std::memory_order_relaxed)))
                    break;
            }
            if (!got) {
                std::unique_lock<std::mutex> l(m_);
                sleepers_.fetch_add(1, std::memory_order_release);
                cv_.wait(l, [&] { return quit_.load() || gen_.load(std::memory_order_acquire) != seen; });
                sleepers_.fetch_sub(1, std::memory_order_release);
                if (quit_.load()) return;
            }
            seen = gen_.load(std::memory_order_acquire);
            slice(i);
            pending_.fetch_sub(1, std::memory_order_release);
        }
    }
    void stop() {
        {
            std::lock_guard<std::mutex> l(m_);
            quit_.store(true);
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
        workers_.clear();
    }
    std::mutex api_, m_;
    std::condition_variable cv_;
    std::vector<std::thread> workers_;
    const CopySeg* segs_ = nullptr;
    int nseg_ = 0;
    size_t total_ = 0, per_ = 0;
    std::atomic<int> pending_{0}, sleepers_{0}, spin_us_{1000};
    std::atomic<unsigned long long> gen_{0};
    std::atomic<bool> quit_{false};
};

static CopyPool& pool() {
    static CopyPool* p = new CopyPool();   // leaked on purpose: worker threads must not be joined from a static destructor
    return *p;
}
static std::atomic<int> g_pool_configured{0};

static int affinity_cpus() {
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof(set), &set) != 0) return 1;
    const int n = CPU_COUNT(&set);
    return n > 0 ? n : 1;
}

static void ensure_pool() {
    if (g_pool_configured.load()) return;
    int want = affinity_cpus() / 2;
    if (want > 8) want = 8;
    if (want < 1) want = 1;
    pool().resize(want - 1);                // the calling thread is the n-th copier
    g_pool_configured.store(1);
}

// "0-15,32-47" -> cpu_set_t
static int parse_cpulist(const char* s, cpu_set_t* set) {
    CPU_ZERO(set);
    int count = 0;
    while (*s) {
        while (*s == ',' || *s == ' ' || *s == '\n') ++s;
        if (!*s) break;
        char* e = nullptr;
        long a = strtol(s, &e, 10);
        if (e == s) break;
        long b = a;
        s = e;
        if (*s == '-') {
            b = strtol(s + 1, &e, 10);
            s = e;
        }
        for (long c = a; c <= b && c < CPU_SETSIZE; ++c) {
            CPU_SET((int)c, set);
            ++count;
        }
    }
    return count;
}

static bool read_small_file(const char* path, char* buf, size_t cap) {
    FILE* f = fopen(path, "r");
    if (!f) return false;
    const size_t n = fread(buf, 1, cap - 1, f);
    fclose(f);
    buf[n] = 0;
    return n > 0;
}

}  // namespace lidar

using namespace lidar;

extern "C" {

int lidar_bind_to_device_numa(int device, int* node_out, int* ncpus_out) {
    if (node_out) *node_out = -1;
    if (ncpus_out) *ncpus_out = affinity_cpus();
    char bdf[32] = "";
    LIDAR_CUDA_TRY(cudaDeviceGetPCIBusId(bdf, sizeof(bdf), device));
    for (char* c = bdf; *c; ++c) *c = (char)tolower(*c);
    char path[128], buf[4096];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bdf);
    if (!read_small_file(path, buf, sizeof(buf))) return LIDAR_OK;     // no sysfs entry: leave everything as it is
    const int node = atoi(buf);
    if (node < 0) return LIDAR_OK;                                     // single-node platform (or a VM that hides it)
    snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
    if (!read_small_file(path, buf, sizeof(buf))) return LIDAR_OK;
    cpu_set_t want, have, both;
    if (parse_cpulist(buf, &want) == 0) return LIDAR_OK;
    CPU_ZERO(&have);
    sched_getaffinity(0, sizeof(have), &have);
    CPU_AND(&both, &want, &have);                                      // never widen a mask the launcher set
    if (CPU_COUNT(&both) == 0) return LIDAR_OK;
    if (sched_setaffinity(0, sizeof(both), &both) != 0) return LIDAR_OK;
    // prefer the node's memory for everything this thread first-touches from now on (MPOL_PREFERRED = 1)
    unsigned long mask[16] = {0};
    if (node < (int)(sizeof(mask) * 8)) {
        mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
        (void)syscall(SYS_set_mempolicy, 1, mask, sizeof(mask) * 8);
    }
    if (node_out) *node_out = node;
    if (ncpus_out) *ncpus_out = CPU_COUNT(&both);
    // the copy pool is re-created inside the new mask
    g_pool_configured.store(0);
    pool().resize(0);
    return LIDAR_OK;
}

int lidar_host_alloc(size_t bytes, void** h_ptr_out) {
    LIDAR_REQUIRE(h_ptr_out != nullptr && bytes > 0, LIDAR_ERR_INVALID, "lidar_host_alloc: bad argument");
    void* p = nullptr;
    LIDAR_CUDA_TRY(cudaHostAlloc(&p, bytes, cudaHostAllocPortable | cudaHostAllocMapped));
    // first touch from this (NUMA-bound) thread; cudaHostAlloc has usually faulted the pages in already
    memset(p, 0, bytes);
    *h_ptr_out = p;
    return LIDAR_OK;
}

int lidar_host_free(void* h_ptr) {
    if (h_ptr) LIDAR_CUDA_TRY(cudaFreeHost(h_ptr));
    return LIDAR_OK;
}

int lidar_copy_async(void* dst, const void* src, size_t bytes, int to_device, void* stream) {
    LIDAR_REQUIRE(bytes == 0 || (dst && src), LIDAR_ERR_INVALID, "lidar_copy_async: NULL buffer");
    if (bytes)
        LIDAR_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost,
                                       as_stream(stream)));
    return LIDAR_OK;
}

int lidar_stream_synchronize(void* stream) {
    LIDAR_CUDA_TRY(cudaStreamSynchronize(as_stream(stream)));
    return LIDAR_OK;
}

int lidar_host_copy_threads(int threads) {
    LIDAR_REQUIRE(threads >= 1 && threads <= 64, LIDAR_ERR_INVALID, "lidar_host_copy_threads: 1..64");
    pool().resize(threads - 1);
    g_pool_configured.store(1);
    return LIDAR_OK;
}

int lidar_host_copy_nontemporal(int on) {
    g_copy_nt.store(on ? 1 : 0);
    return LIDAR_OK;
}

int lidar_host_memcpy(void* dst, const void* src, size_t bytes) {
    LIDAR_REQUIRE(bytes == 0 || (dst && src), LIDAR_ERR_INVALID, "lidar_host_memcpy: NULL buffer");
    if (bytes == 0) return LIDAR_OK;
    ensure_pool();
    const CopySeg seg{static_cast<char*>(dst), static_cast<const char*>(src), bytes};
    pool().copy(&seg, 1);
    return LIDAR_OK;
}

int lidar_host_copy_wake(void) {
    ensure_pool();
    pool().wake();
    return LIDAR_OK;
}

int lidar_host_memcpy_batch(int count, void* const* dst, const void* const* src, const size_t* bytes) {
    LIDAR_REQUIRE(count >= 0 && count <= 64 && (count == 0 || (dst && src && bytes)), LIDAR_ERR_INVALID,
                  "lidar_host_memcpy_batch: bad argument (at most 64 segments)");
    CopySeg segs[64];
    for (int k = 0; k < count; ++k) {
        LIDAR_REQUIRE(bytes[k] == 0 || (dst[k] && src[k]), LIDAR_ERR_INVALID, "lidar_host_memcpy_batch: NULL segment %d", k);
        segs[k] = CopySeg{static_cast<char*>(dst[k]), static_cast<const char*>(src[k]), bytes[k]};
    }
    if (count == 0) return LIDAR_OK;
    ensure_pool();
    pool().copy(segs, count);
    return LIDAR_OK;
}

}  // extern "C"
