// dev_host.hpp — TEST INFRASTRUCTURE: a device policy that runs the kernel bodies of
// gym-fish_b200/csrc/*.cuh in plain CPU loops (block by block, thread by thread).
// It exists so that AA-pattern indexing, bounce-back, z-face ops and the IB kernels can be checked
// against the fp64 oracle in this GPU-less container before GPU time is spent.  It is never loaded by
// the product (gym-fish_b200/_abi.py knows only the CUDA library); results from it are not benchmarks.
#pragma once
#include <atomic>
#include <chrono>
#include <thread>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unistd.h>

namespace fg {

struct PeerBlob;

class HostDev {
public:
    std::string err;
    long long launches = 0;
    int wait_ms = 300;   // how long a slab waits for its neighbour's halo before reporting it is behind

    bool init(int, std::string &) { return true; }
    void enter() {}
    void shutdown() {}
    void *alloc(size_t bytes, std::string &e) {
        void *p = std::malloc(bytes ? bytes : 1);
        if (!p) e = "host emulation: malloc failed";
        return p;
    }
    void free(void *p) { std::free(p); }
    bool h2d(void *d, const void *s, size_t n) { std::memcpy(d, s, n); return true; }
    bool d2h(void *d, const void *s, size_t n) { std::memcpy(d, s, n); return true; }
    bool zero(void *d, size_t n) { std::memset(d, 0, n); return true; }
    bool sync() { return true; }
    void *alloc_host(size_t bytes, std::string &e) { return alloc(bytes, e); }
    void free_host(void *p) { std::free(p); }
    bool h2d_async(void *d, const void *s, size_t n) { return h2d(d, s, n); }
    bool d2h_async(void *d, const void *s, size_t n) { return d2h(d, s, n); }
    bool ev_record(int) { return true; }
    bool ev_sync(int) { return true; }
    void tic() { t0_ = std::chrono::steady_clock::now(); }
    void marks_reset() {}
    void mark(int) {}
    double marks_elapsed(int) { return 0.0; }
    void toc_record() { t1_ = std::chrono::steady_clock::now(); }
    double toc_elapsed(bool, bool &ok) { ok = true; return std::chrono::duration<double, std::milli>(t1_ - t0_).count(); }

    template <class K, class P>
    bool launch(Dim3 g, const P &p) {
        ++launches;
        for (int bz = 0; bz < g.z; ++bz)
            for (int by = 0; by < g.y; ++by)
                for (int bx = 0; bx < g.x; ++bx)
                    for (int tx = 0; tx < K::kThreads; ++tx) K::run(p, bx, by, bz, tx);
        return true;
    }

    // block-phased kernels: CTA barrier between the phases -> per block: all threads of phase 0, then phase 1, ...
    template <class K, class P>
    bool launch_block_phased(Dim3 g, const P &p) {
        ++launches;
        for (int bz = 0; bz < g.z; ++bz)
            for (int by = 0; by < g.y; ++by)
                for (int bx = 0; bx < g.x; ++bx)
                    for (int ph = 0; ph < K::kBlockPhases; ++ph)
                        for (int tx = 0; tx < K::kThreads; ++tx) K::run(p, bx, by, bz, tx, ph);
        return true;
    }

    // ticket-scheduled kernels: on the GPU the phases interleave along a wavefront; here every CTA of phase 0, then phase 1
    template <class K, class P>
    bool launch_ticketed(Dim3 g, const P &p) {
        ++launches;
        for (int ph = 0; ph < K::kGridPhases; ++ph)
            for (int bz = 0; bz < g.z; ++bz)
                for (int by = 0; by < g.y; ++by)
                    for (int bx = 0; bx < g.x; ++bx)
                        for (int tx = 0; tx < K::kThreads; ++tx) K::run(p, bx, by, bz, tx, ph);
        return true;
    }
    bool zero_on_current(void *d, size_t n) { std::memset(d, 0, n); return true; }

    // phased kernels (one cooperative launch on the GPU): phases in order, every item of a phase before the next
    bool supports_phased() const { return true; }
    template <class K, class P>
    bool launch_phased(long long, const P &p) {
        ++launches;
        for (int ph = 0; ph < K::kPhases; ++ph) {
            const long long n = K::items(p, ph);
            for (long long i = 0; i < n; ++i) K::item(p, ph, i);
        }
        return true;
    }

    // peers: only same-process handles (raw pointers) — lets the CPU tests run two slabs in one process
    template <class Blob>
    bool export_peer(pop_t *f, int *flags, void *x, Blob &b, std::string &) {
        b.pid = int(getpid()); b.device = -1;
        b.f_ptr = reinterpret_cast<uint64_t>(f); b.flag_ptr = reinterpret_cast<uint64_t>(flags); b.x_ptr = reinterpret_cast<uint64_t>(x);
        return true;
    }
    template <class Blob>
    bool open_peer(const Blob &b, pop_t **f, int **flags, void **x, std::string &e) {
        if (b.pid != int(getpid())) { e = "host emulation: peers must live in the same process"; return false; }
        *f = reinterpret_cast<pop_t *>(b.f_ptr); *flags = reinterpret_cast<int *>(b.flag_ptr); *x = reinterpret_cast<void *>(b.x_ptr);
        return true;
    }
    // exchange counters (IB across slabs): ranks run in separate host threads in the tests, so these really wait
    bool signal_counters(int *mine, int *const *targets, int n, bool bump) {
        const int v = mine[0] + 1;
        std::atomic_thread_fence(std::memory_order_release);
        for (int i = 0; i < n; ++i)
            if (targets[i]) *static_cast<volatile int *>(targets[i]) = v;
        if (bump) mine[0] = v;
        return true;
    }
    bool wait_counters(int *mine, int *const *sources, int n) {
        const int v = mine[0] + 1;
        const auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < n; ++i) {
            if (!sources[i]) continue;
            while (*static_cast<volatile int *>(sources[i]) < v) {
                if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(20)) {
                    err = "host emulation: timed out waiting for another rank's exchange";
                    return false;
                }
                std::this_thread::yield();
            }
        }
        std::atomic_thread_fence(std::memory_order_acquire);
        return true;
    }
    void close_peers() {}
    int peer_timeout() const { return 0; }      // the emulated waits report a neighbour that is behind directly
    int current() const { return cur_; }   // streams: everything runs sequentially here, in submission order
    bool fork_to(int s) { cur_ = s; return true; }
    bool switch_to(int s) { cur_ = s; return true; }
    bool join_from(int) { return true; }
    bool graph_begin(const GraphKey &) { return false; }   // no graphs in the emulation: everything runs directly
    bool graph_end() { return true; }
    void graph_abort() {}
    void graph_clear() {}
    bool signal_flags(int *mine, int *lo, int *hi) {
        const int value = ++mine[2];
        if (lo) *lo = value;
        if (hi) *hi = value;
        return true;
    }
    bool wait_flags(const int *flags, bool lo, bool hi) {
        const int value = flags[2];
        const auto t0 = std::chrono::steady_clock::now();
        const volatile int *vf = flags;
        while ((lo && vf[0] < value) || (hi && vf[1] < value)) {
            if (std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(wait_ms)) {
                err = "host emulation: neighbour slab is behind (step the slabs in lock-step or from separate threads)";
                return false;
            }
            std::this_thread::yield();
        }
        std::atomic_thread_fence(std::memory_order_acquire);
        return true;
    }

private:
    int cur_ = 0;
    std::chrono::steady_clock::time_point t0_, t1_;
};

}  // namespace fg
