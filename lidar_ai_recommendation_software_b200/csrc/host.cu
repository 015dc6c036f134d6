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
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace lidar {

// ---- worker pool ---------------------------------------------------------------------------------
class CopyPool {
  public:
    ~CopyPool() { stop(); }
    void resize(int n) {
        std::lock_guard<std::mutex> g(api_);
        if (n == (int)workers_.size()) return;
        stop();
        quit_ = false;
        gen_ = 0;
        for (int i = 0; i < n; ++i) workers_.emplace_back([this, i] { run(i); });
    }
    int size() {
        std::lock_guard<std::mutex> g(api_);
        return (int)workers_.size();
    }
    // copy [src, src + bytes) -> dst with the pool + the calling thread
    void copy(char* dst, const char* src, size_t bytes) {
        std::lock_guard<std::mutex> g(api_);
        const size_t parts = workers_.size() + 1;
        if (workers_.empty() || bytes < (size_t)(1 << 20)) {
            memcpy(dst, src, bytes);
            return;
        }
        // 64-byte aligned slices
        const size_t per = ((bytes + parts - 1) / parts + 63) & ~(size_t)63;
        {
            std::lock_guard<std::mutex> l(m_);
            dst_ = dst; src_ = src; bytes_ = bytes; per_ = per;
            pending_ = (int)workers_.size();
            ++gen_;
        }
        cv_.notify_all();
        slice((int)workers_.size());            // the caller takes the last slice
        std::unique_lock<std::mutex> l(m_);
        done_.wait(l, [this] { return pending_ == 0; });
    }

  private:
    void slice(int i) {
        const size_t a = (size_t)i * per_;
        if (a >= bytes_) return;
        const size_t len = a + per_ <= bytes_ ? per_ : bytes_ - a;
        memcpy(dst_ + a, src_ + a, len);
    }
    void run(int i) {
        unsigned long long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [&] { return quit_ || gen_ != seen; });
                if (quit_) return;
                seen = gen_;
            }
            slice(i);
            {
                std::lock_guard<std::mutex> l(m_);
                if (--pending_ == 0) done_.notify_one();
            }
        }
    }
    void stop() {
        {
            std::lock_guard<std::mutex> l(m_);
            quit_ = true;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
        workers_.clear();
    }
    std::mutex api_, m_;
    std::condition_variable cv_, done_;
    std::vector<std::thread> workers_;
    char* dst_ = nullptr;
    const char* src_ = nullptr;
    size_t bytes_ = 0, per_ = 0;
    int pending_ = 0;
    unsigned long long gen_ = 0;
    bool quit_ = false;
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

int lidar_host_copy_threads(int threads) {
    LIDAR_REQUIRE(threads >= 1 && threads <= 64, LIDAR_ERR_INVALID, "lidar_host_copy_threads: 1..64");
    pool().resize(threads - 1);
    g_pool_configured.store(1);
    return LIDAR_OK;
}

int lidar_host_memcpy(void* dst, const void* src, size_t bytes) {
    LIDAR_REQUIRE(bytes == 0 || (dst && src), LIDAR_ERR_INVALID, "lidar_host_memcpy: NULL buffer");
    if (bytes == 0) return LIDAR_OK;
    ensure_pool();
    pool().copy(static_cast<char*>(dst), static_cast<const char*>(src), bytes);
    return LIDAR_OK;
}

}  // extern "C"
