// dev_cuda.cuh — the product's device policy: CUDA runtime only (no torch, no NCCL on the sim path).
// One non-blocking stream per handle, CUDA events for timing, CUDA-IPC / peer access for z-neighbours,
// acquire/release flags in peer memory for the per-step neighbour hand-shake.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unistd.h>
#include <map>
#include <utility>
#include <vector>

#include "lbm_core.cuh"

namespace fg {

template <class K, class P>
__global__ void __launch_bounds__(K::kThreads, K::kMinBlocks) kern(const __grid_constant__ P p) {
    K::run(p, int(blockIdx.x), int(blockIdx.y), int(blockIdx.z), int(threadIdx.x));
}

// CTA-phased kernel: K::cta() holds its own CTA barrier(s) (the host emulation runs K::run(..., phase) per phase instead)
template <class K, class P>
__global__ void __launch_bounds__(K::kThreads, K::kMinBlocks) kern_bp(const __grid_constant__ P p) {
    K::cta(p, int(blockIdx.x), int(threadIdx.x));
}

// ticket-scheduled kernel (StreamCollidePair): one-dimensional grid, the CTA finds its work through an atomic ticket
template <class K, class P>
__global__ void __launch_bounds__(K::kThreads, K::kMinBlocks) kern_ticket(const __grid_constant__ P p) {
    K::cta(p, int(threadIdx.x));
}

// phased kernel: K::kPhases phases of grid-stride work separated by grid-wide barriers (cooperative launch)
template <class K, class P>
__global__ void __launch_bounds__(K::kThreads, K::kMinBlocks) kern_phased(const __grid_constant__ P p) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    for (int ph = 0; ph < K::kPhases; ++ph) {
        const long long n = K::items(p, ph);
        for (long long i = (long long)blockIdx.x * K::kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * K::kThreads)
            K::item(p, ph, i);
        if (ph + 1 < K::kPhases) grid.sync();
    }
}

// neighbour hand-shake: one thread, system-scope release / acquire on a word in (peer) device memory
__global__ void signal_kernel(int *mine, int *lo, int *hi) {
    const int value = mine[2] + 1;   // my halo counter lives on the device so that the launch arguments never change
    mine[2] = value;
    __threadfence_system();
    if (lo) asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(lo), "r"(value) : "memory");
    if (hi) asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(hi), "r"(value) : "memory");
}
__global__ void wait_kernel(int *flags, int lo, int hi, unsigned long long timeout_ns, volatile int *host_word) {
    const int value = flags[2];      // neighbours must have pushed as many halos as I have
    unsigned long long t0;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    for (int s = 0; s < 2; ++s) {
        if (!(s == 0 ? lo : hi)) continue;
        for (;;) {
            int v;
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flags + s) : "memory");
            if (v >= value) break;
            unsigned long long t;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            if (t - t0 > timeout_ns) { flags[3] = 1; if (host_word) *host_word = 1; return; }   // sticky; the host word (pinned) lets fg_step see it without a sync
            __nanosleep(200);
        }
    }
}

// exchange counters between ranks (IB across slabs): value = my counter + 1
struct CounterList { int *p[8]; };
__global__ void signal_counters_kernel(int *mine, CounterList t, int bump) {
    const int v = mine[0] + 1;
    __threadfence_system();
    for (int i = 0; i < 8; ++i)
        if (t.p[i]) asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(t.p[i]), "r"(v) : "memory");
    if (bump) mine[0] = v;
}
__global__ void wait_counters_kernel(int *mine, CounterList s, unsigned long long timeout_ns, volatile int *host_word) {
    const int v = mine[0] + 1;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    for (int i = 0; i < 8; ++i) {
        if (!s.p[i]) continue;
        for (;;) {
            int x;
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(x) : "l"(s.p[i]) : "memory");
            if (x >= v) break;
            unsigned long long t;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            if (t - t0 > timeout_ns) { mine[3] = 1; if (host_word) *host_word = 2; return; }
            __nanosleep(200);
        }
    }
}

// Every kernel this library launches registers itself here (static initialisation of KernelReg<...>::done, forced by the
// launch wrappers), and CudaDev::init loads them all before the first step.  Under CUDA's lazy module loading a kernel is
// loaded at its first launch, which synchronises the context — behind a neighbour-wait kernel that may be spinning for exactly
// that launch (peered handles in one process: FG_EPEER in the first cases of a run, gpu passes b24 - b26; one process per GPU:
// a stall in the first steps).  Loading our own kernels leaves the application's modules (CUDA_MODULE_LOADING) alone.
// Internal linkage throughout (anonymous namespace): the fp32 and the 16-bit build are two shared libraries with the same symbol
// names, and inline statics / static members of templates are GNU-unique symbols that the dynamic linker would merge across them.
namespace {
std::vector<const void *> &kernel_registry() {
    static std::vector<const void *> v;
    return v;
}
bool register_kernel(const void *f) {
    kernel_registry().push_back(f);
    return true;
}
template <class K, class P, int FORM>
struct KernelReg {
    static const bool done;
};
template <class K, class P, int FORM>
const bool KernelReg<K, P, FORM>::done =
    register_kernel(FORM == 0 ? reinterpret_cast<const void *>(kern<K, P>) : FORM == 1 ? reinterpret_cast<const void *>(kern_bp<K, P>)
                                                                                      : reinterpret_cast<const void *>(kern_phased<K, P>));
}  // namespace

class CudaDev {
public:
    std::string err;
    long long launches = 0;
    long long graph_launches = 0;      // cudaGraphLaunch calls (FgStats.graph_launches)

    int sm_count() const { return sm_count_; }
    // IbInterpSpread: one warp per marker on the GPU (4 per CTA); the host emulation runs the per-thread form (2 markers per block)
    static int interp_spread_blocks(int n_markers) { return (n_markers + 3) / 4; }
    // NVTX ranges per kernel class (step / ib / collide / faces): free when no tool is attached (nvtx3 is header-only and
    // resolves its injection library lazily), and they name the phases of a substep in an nsys / ncu timeline
    void range_push(const char *name) { nvtxRangePushA(name); }
    void range_pop() { nvtxRangePop(); }

    bool init(int device, std::string &e) {
        int n = 0;
        // A handle owns 6 streams + a copy stream.  The default of 8 hardware work queues per device makes streams share queues, and
        // a queue is served in submission order: harmless for one handle per process (the production layout), but with several
        // PEERED handles in one process (tests, smoke) a neighbour-wait kernel could sit in front of the very push it waits for.
        // More queues (the maximum is 32) before the context exists; a value the application has set itself is left alone.
        setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
        cudaError_t rc = cudaGetDeviceCount(&n);
        if (rc != cudaSuccess || n == 0) {
            e = std::string("no CUDA device: ") + cudaGetErrorString(rc) + " — this library has no CPU path";
            return false;
        }
        if (device < 0 || device >= n) { e = "FgConfig.device out of range"; return false; }
        device_ = device;
        if (!ck(cudaSetDevice(device_), "cudaSetDevice")) { e = err; return false; }
        cudaDeviceProp prop;
        if (!ck(cudaGetDeviceProperties(&prop, device_), "cudaGetDeviceProperties")) { e = err; return false; }
        if (prop.major < 10) {
            e = std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                "; this library is built for sm_100a only";
            return false;
        }
        sm_count_ = prop.multiProcessorCount;
        coop_ = prop.cooperativeLaunch != 0;
        // Streams 0/1 (main + its thin branch) run at the highest priority, 2/3 (the far-plane collide that overlaps
        // the IB kernels, sim.hpp step()) at the lowest: pending CTAs of the short IB kernels are then dispatched
        // ahead of the tens of thousands of queued collide CTAs instead of behind them.
        int least = 0, greatest = 0;
        cudaDeviceGetStreamPriorityRange(&least, &greatest);
        prio_low_ = least; prio_high_ = greatest;
        debug_sync_ = getenv("FG_DEBUG_SYNC") != nullptr;       // debugging aid: serialise every launch
        // the neighbour-wait kernels give up after this long (ranks out of step); FG_PEER_TIMEOUT_MS shortens it for tests
        if (const char *t = getenv("FG_PEER_TIMEOUT_MS")) {
            const long long ms = atoll(t);
            if (ms > 0) peer_timeout_ns_ = (unsigned long long)ms * 1000000ull;
        }
        // one pinned, device-mapped word the wait kernels raise when they time out: fg_step reads it without a sync
        {
            int *w = nullptr;
            if (!ck(cudaHostAlloc(reinterpret_cast<void **>(&w), sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable), "cudaHostAlloc(timeout word)")) { e = err; return false; }
            *w = 0;
            timeout_word_ = w;
        }
        bool ok = true;
        for (int i = 0; i < kStreams && ok; ++i)
            ok = ck(cudaStreamCreateWithPriority(&s_[i], cudaStreamNonBlocking, (i < 2 || i >= 4) ? greatest : least), "cudaStreamCreate") &&
                 ck(cudaEventCreateWithFlags(&fork_ev_[i], cudaEventDisableTiming), "cudaEventCreate") &&
                 ck(cudaEventCreateWithFlags(&join_ev_[i], cudaEventDisableTiming), "cudaEventCreate");
        stream_ = s_[0];
        ok = ok && ck(cudaStreamCreateWithFlags(&copy_s_, cudaStreamNonBlocking), "cudaStreamCreate(copy)");
        {   // load every kernel of the library now (see kernel_registry)
            cudaFuncAttributes fa;
            for (const void *f : kernel_registry()) if (ok) ok = ck(cudaFuncGetAttributes(&fa, f), "kernel preload");
            const void *small[] = {reinterpret_cast<const void *>(signal_kernel), reinterpret_cast<const void *>(wait_kernel),
                                   reinterpret_cast<const void *>(signal_counters_kernel), reinterpret_cast<const void *>(wait_counters_kernel)};
            for (const void *f : small) if (ok) ok = ck(cudaFuncGetAttributes(&fa, f), "kernel preload");
        }
        for (int i = 0; i < 2 && ok; ++i)
            ok = ck(cudaEventCreate(&tev_[i][0]), "cudaEventCreate") && ck(cudaEventCreate(&tev_[i][1]), "cudaEventCreate");
        if (!ok) e = err;
        return ok;
    }
    void shutdown() {
        if (device_ < 0) return;
        cudaSetDevice(device_);
        graph_clear();
        for (int i = 0; i < kStreams; ++i) {
            if (s_[i]) { cudaStreamSynchronize(s_[i]); cudaStreamDestroy(s_[i]); s_[i] = nullptr; }
            if (fork_ev_[i]) cudaEventDestroy(fork_ev_[i]);
            if (join_ev_[i]) cudaEventDestroy(join_ev_[i]);
            fork_ev_[i] = join_ev_[i] = nullptr;
        }
        stream_ = nullptr;
        if (copy_s_) { cudaStreamSynchronize(copy_s_); cudaStreamDestroy(copy_s_); copy_s_ = nullptr; }
        for (auto &pr : tev_)
            for (auto &e : pr) { if (e) cudaEventDestroy(e); e = nullptr; }
        if (timeout_word_) { cudaFreeHost(const_cast<int *>(timeout_word_)); timeout_word_ = nullptr; }
        for (auto &pool : marks_) { for (auto e : pool) cudaEventDestroy(e); pool.clear(); }
        for (auto &e : named_) { if (e) cudaEventDestroy(e); e = nullptr; }
    }

    // Every extern "C" entry point calls enter() once (abi_impl.hpp FG_TRY): the methods below assume the handle's device
    // is current on the calling thread and do not set it again (33 cudaSetDevice calls per step before).
    void enter() { if (device_ >= 0) cudaSetDevice(device_); }

    void *alloc(size_t bytes, std::string &e) {
        void *p = nullptr;
        cudaError_t rc = cudaMalloc(&p, bytes ? bytes : 4);
        if (rc != cudaSuccess) {
            e = std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(rc);
            cudaGetLastError();
            return nullptr;
        }
        return p;
    }
    void free(void *p) {
        if (!p) return;
        cudaFree(p);
    }
    bool h2d(void *d, const void *s, size_t n) {
        return ck(cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, stream_), "H2D") && ck(cudaStreamSynchronize(stream_), "H2D sync");
    }
    bool d2h(void *d, const void *s, size_t n) {
        return ck(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, stream_), "D2H") && ck(cudaStreamSynchronize(stream_), "D2H sync");
    }
    bool zero(void *d, size_t n) {
        if (gmode_ == 2) return true;   // replaying: the memset is a node of the graph
        return ck(cudaMemsetAsync(d, 0, n, stream_), "memset");
    }
    bool sync() {
        return ck(cudaStreamSynchronize(stream_), "stream sync");
    }
    // pinned host memory + asynchronous copies on the handle's stream + a few named events (IbState)
    void *alloc_host(size_t bytes, std::string &e) {
        void *p = nullptr;
        cudaError_t rc = cudaMallocHost(&p, bytes ? bytes : 4);
        if (rc != cudaSuccess) {
            e = std::string("cudaMallocHost(") + std::to_string(bytes) + "): " + cudaGetErrorString(rc);
            cudaGetLastError();
            return nullptr;
        }
        return p;
    }
    void free_host(void *p) {
        if (!p) return;
        cudaFreeHost(p);
    }
    bool h2d_async(void *d, const void *pinned, size_t n) {
        if (gmode_ == 2) return true;
        return ck(cudaMemcpyAsync(d, pinned, n, cudaMemcpyHostToDevice, stream_), "H2D async");
    }
    bool d2h_async(void *pinned, const void *s, size_t n) {
        if (gmode_ == 2) return true;
        return ck(cudaMemcpyAsync(pinned, s, n, cudaMemcpyDeviceToHost, stream_), "D2H async");
    }
    bool ev_record(int id) {
        if (id < 0 || id >= kNamedEvents) return false;
        if (!named_[id] && !ck(cudaEventCreateWithFlags(&named_[id], cudaEventDisableTiming), "cudaEventCreate")) return false;
        if (gmode_ == 2) return true;
        // inside a capture the record must become an event-record NODE the host can wait on, not a capture-internal edge
        return ck(cudaEventRecordWithFlags(named_[id], stream_, gmode_ == 1 ? cudaEventRecordExternal : cudaEventRecordDefault),
                  "cudaEventRecord");
    }
    bool ev_sync(int id) {
        if (id < 0 || id >= kNamedEvents || !named_[id]) return true;
        return ck(cudaEventSynchronize(named_[id]), "cudaEventSynchronize");
    }
    // Host -> device copy that starts NOW on the copy stream (beside whatever the handle's streams are running), after the
    // work event `after` marks (the last reader of the destination); event `done` is recorded behind it and stream 0 waits
    // for it, so everything queued on the handle's streams from here on sees the data.  Never called inside a capture.
    bool upload_early(void *d, const void *pinned, size_t n, int done, int after) {
        if (done < 0 || done >= kNamedEvents || gmode_ != 0) { err = "upload_early: bad event or called inside a graph capture"; return false; }
        if (!named_[done] && !ck(cudaEventCreateWithFlags(&named_[done], cudaEventDisableTiming), "cudaEventCreate")) return false;
        if (after >= 0 && after < kNamedEvents && named_[after] && !ck(cudaStreamWaitEvent(copy_s_, named_[after], 0), "copy stream wait")) return false;
        return ck(cudaMemcpyAsync(d, pinned, n, cudaMemcpyHostToDevice, copy_s_), "H2D early") &&
               ck(cudaEventRecord(named_[done], copy_s_), "cudaEventRecord(copy)") && ck(cudaStreamWaitEvent(stream_, named_[done], 0), "wait for upload");
    }
    // Timing of one fg_step call: two event pairs used alternately, so that a call can record its own pair while the
    // previous call (which returned before its collide finished) still owns the other one.
    void tic() {
        tpair_ ^= 1;
        cudaEventRecord(tev_[tpair_][0], stream_);
    }
    void toc_record() { cudaEventRecord(tev_[tpair_][1], stream_); }
    // elapsed ms of the latest recorded pair; wait == false: only if it has completed (ok tells)
    double toc_elapsed(bool wait, bool &ok) {
        ok = false;
        cudaEvent_t e1 = tev_[tpair_][1];
        if (wait) { if (cudaEventSynchronize(e1) != cudaSuccess) return 0.0; }
        else if (cudaEventQuery(e1) != cudaSuccess) { cudaGetLastError(); return 0.0; }
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, tev_[tpair_][0], e1) != cudaSuccess) { cudaGetLastError(); return 0.0; }
        ok = true;
        return double(ms);
    }

    // FG_FLAG_PROFILE: event pairs around launches of one kernel class (0 = stream-collide, 1 = immersed boundary);
    // consecutive mark(c) calls open / close an interval on the launching stream
    void marks_reset() { nmarks_[0] = nmarks_[1] = 0; }
    void mark(int c) {
        auto &pool = marks_[c];
        if (nmarks_[c] >= kMaxMarks) return;   // long runs: the first kMaxMarks/2 intervals are a fair sample
        if (int(pool.size()) <= nmarks_[c]) {
            cudaEvent_t e;
            if (cudaEventCreate(&e) != cudaSuccess) return;
            pool.push_back(e);
        }
        cudaEventRecord(pool[nmarks_[c]++], stream_);
    }
    double marks_elapsed(int c) {
        double tot = 0;
        for (int i = 0; i + 1 < nmarks_[c]; i += 2) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, marks_[c][i], marks_[c][i + 1]) == cudaSuccess) tot += ms;
        }
        return tot;
    }
    int marks_intervals(int c) const { return nmarks_[c] / 2; }

    template <class K, class P>
    bool launch(Dim3 g, const P &p) {
        ++launches;
        if (gmode_ == 2) return true;   // replaying a captured substep
        (void)KernelReg<K, P, 0>::done;
        return launch_on_current(kern<K, P>, dim3(g.x, g.y, g.z), K::kThreads, p);
    }
    // The priority travels as a launch attribute, not only as a stream property: a captured kernel node keeps it,
    // whereas nodes of a graph launched into one stream would otherwise all run at that stream's priority.
    template <class P>
    bool launch_on_current(void (*kernel)(const P), dim3 grid, int threads, const P &p) {
        cudaLaunchConfig_t lc{};
        lc.gridDim = grid; lc.blockDim = dim3(threads); lc.dynamicSmemBytes = 0; lc.stream = s_[cur_];
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributePriority;
        at[0].val.priority = (cur_ < 2 || cur_ >= 4) ? prio_high_ : prio_low_;
        lc.attrs = at; lc.numAttrs = 1;
        if (!ck(cudaLaunchKernelEx(&lc, kernel, p), "kernel launch")) return false;
        if (debug_sync_ && gmode_ == 0) return ck(cudaDeviceSynchronize(), "debug sync");
        return true;
    }
    // Kernel launches go to stream `current()`.  fork_to(s): stream s starts where the current stream is now and
    // becomes current; switch_to(s): make s current (work already queued elsewhere keeps running beside it);
    // join_from(s): the current stream waits for everything queued on s.  Inside a capture these are the parallel
    // branches of the graph.  Copies, memsets, event records and the neighbour WAITS always use stream 0; the halo signal follows
    // the current stream.
    int current() const { return cur_; }
    bool fork_to(int s) {
        const int from = cur_;
        cur_ = s;
        if (gmode_ == 2) return true;
        return ck(cudaEventRecord(fork_ev_[s], s_[from]), "fork record") && ck(cudaStreamWaitEvent(s_[s], fork_ev_[s], 0), "fork wait");
    }
    bool switch_to(int s) { cur_ = s; return true; }
    bool join_from(int s) {
        if (gmode_ == 2) return true;
        return ck(cudaEventRecord(join_ev_[s], s_[s]), "join record") && ck(cudaStreamWaitEvent(s_[cur_], join_ev_[s], 0), "join wait");
    }

    template <class K, class P>
    bool launch_block_phased(Dim3 g, const P &p) {
        ++launches;
        if (gmode_ == 2) return true;
        (void)KernelReg<K, P, 1>::done;
        return launch_on_current(kern_bp<K, P>, dim3(g.x, g.y, g.z), K::kThreads, p);
    }
    // g = (x-blocks, rows, planes) of ONE phase; the launch holds K::kGridPhases times as many CTAs in one dimension
    template <class K, class P>
    bool launch_ticketed(Dim3 g, const P &p) {
        ++launches;
        if (gmode_ == 2) return true;
        const long long n = (long long)g.x * g.y * g.z * K::kGridPhases;
        if (n > 0x7fffffffll) { err = "ticketed launch too large"; return false; }
        return launch_on_current(kern_ticket<K, P>, dim3(unsigned(n)), K::kThreads, p);
    }
    bool zero_on_current(void *d, size_t n) {
        if (gmode_ == 2) return true;
        return ck(cudaMemsetAsync(d, 0, n, s_[cur_]), "memset");
    }
    bool supports_phased() const { return coop_; }
    template <class K, class P>
    bool launch_phased(long long max_items, const P &p) {
        ++launches;
        if (gmode_ == 2) return true;
        static thread_local int per_sm = 0;      // per kernel instantiation
        if (per_sm == 0) {
            (void)KernelReg<K, P, 2>::done;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern_phased<K, P>, K::kThreads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
        }
        long long want = (max_items + K::kThreads - 1) / K::kThreads;
        const long long cap = (long long)per_sm * sm_count_;
        if (want > cap) want = cap;
        if (want < 1) want = 1;
        void *args[] = {const_cast<P *>(&p)};
        return ck(cudaLaunchCooperativeKernel(reinterpret_cast<void *>(kern_phased<K, P>), dim3(unsigned(want)), dim3(K::kThreads), args, 0, stream_),
                  "cooperative launch");
    }

    // ---- z-neighbour lattices: same process => raw pointer (+ peer access); other process => CUDA IPC
    template <class Blob>
    bool export_peer(pop_t *f, int *flags, void *x, Blob &b, std::string &e) {
        b.pid = int(getpid()); b.device = device_;
        b.f_ptr = reinterpret_cast<uint64_t>(f); b.flag_ptr = reinterpret_cast<uint64_t>(flags); b.x_ptr = reinterpret_cast<uint64_t>(x);
        cudaIpcMemHandle_t h1, h2, h3;
        static_assert(sizeof(h1) == 64, "cudaIpcMemHandle_t is 64 bytes");
        if (!ck(cudaIpcGetMemHandle(&h1, f), "cudaIpcGetMemHandle(lattice)") || !ck(cudaIpcGetMemHandle(&h2, flags), "cudaIpcGetMemHandle(flags)")) {
            e = err;
            return false;
        }
        std::memcpy(b.f_ipc, &h1, 64); std::memcpy(b.flag_ipc, &h2, 64);
        if (x) {
            if (!ck(cudaIpcGetMemHandle(&h3, x), "cudaIpcGetMemHandle(exchange)")) { e = err; return false; }
            std::memcpy(b.x_ipc, &h3, 64);
        }
        return true;
    }
    bool signal_counters(int *mine, int *const *targets, int n, bool bump) {
        ++launches;
        if (gmode_ == 2) return true;
        CounterList t{};
        for (int i = 0; i < n && i < 8; ++i) t.p[i] = targets ? targets[i] : nullptr;
        signal_counters_kernel<<<1, 1, 0, stream_>>>(mine, t, bump ? 1 : 0);
        return ck(cudaGetLastError(), "signal launch");
    }
    bool wait_counters(int *mine, int *const *sources, int n) {
        ++launches;
        if (gmode_ == 2) return true;
        CounterList s{};
        for (int i = 0; i < n && i < 8; ++i) s.p[i] = sources[i];
        wait_counters_kernel<<<1, 1, 0, stream_>>>(mine, s, peer_timeout_ns_, timeout_word_);
        return ck(cudaGetLastError(), "wait launch");
    }
    template <class Blob>
    bool open_peer(const Blob &b, pop_t **f, int **flags, void **x, std::string &e) {
        if (b.pid == int(getpid())) {
            if (b.device != device_) {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, device_, b.device);
                if (!can) { e = "devices " + std::to_string(device_) + " and " + std::to_string(b.device) + " have no peer access"; return false; }
                cudaError_t rc = cudaDeviceEnablePeerAccess(b.device, 0);
                if (rc != cudaSuccess && rc != cudaErrorPeerAccessAlreadyEnabled) { ck(rc, "cudaDeviceEnablePeerAccess"); e = err; return false; }
                cudaGetLastError();
            }
            *f = reinterpret_cast<pop_t *>(b.f_ptr); *flags = reinterpret_cast<int *>(b.flag_ptr); *x = reinterpret_cast<void *>(b.x_ptr);
            return true;
        }
        void *pf = open_ipc(b.f_ipc), *pg = open_ipc(b.flag_ipc);
        if (!pf || !pg) { e = err; return false; }
        *f = static_cast<pop_t *>(pf); *flags = static_cast<int *>(pg);
        *x = nullptr;
        if (b.x_ptr) {
            *x = open_ipc(b.x_ipc);
            if (!*x) { e = err; return false; }
        }
        return true;
    }
    void close_peers() {
        if (device_ < 0) return;
        for (auto &o : opened_) cudaIpcCloseMemHandle(o.second);
        opened_.clear();
    }
    bool signal_flags(int *mine, int *lo, int *hi) {
        ++launches;
        if (gmode_ == 2) return true;
        signal_kernel<<<1, 1, 0, s_[cur_]>>>(mine, lo, hi);        // on the current stream: the halo branch signals as soon as ITS push is queued
        return ck(cudaGetLastError(), "signal launch");
    }
    bool wait_flags(int *flags, bool lo, bool hi) {
        if (!lo && !hi) return true;
        ++launches;
        if (gmode_ == 2) return true;
        wait_kernel<<<1, 1, 0, stream_>>>(flags, lo, hi, peer_timeout_ns_, timeout_word_);
        return ck(cudaGetLastError(), "wait launch");
    }

    // ---- per-substep CUDA graphs: capture the stream once per key, replay afterwards.  While replaying (gmode_ 2)
    // launches and async copies are not enqueued — the caller's host code still runs and keeps its state in step.
    bool graph_begin(const GraphKey &key) {
        auto it = graphs_.find(key);
        if (it != graphs_.end()) { gmode_ = 2; gexec_ = it->second; return true; }
        if (graphs_.size() >= 64) graph_clear();
        if (cudaStreamBeginCapture(stream_, cudaStreamCaptureModeRelaxed) != cudaSuccess) { cudaGetLastError(); return false; }
        gmode_ = 1; gkey_ = key; cur_ = 0;
        return true;
    }
    bool graph_end() {
        if (gmode_ == 1) {
            gmode_ = 0;
            cudaGraph_t g = nullptr;
            if (!ck(cudaStreamEndCapture(stream_, &g), "cudaStreamEndCapture")) return false;
            cudaGraphExec_t e = nullptr;
            // per-node priorities (the launch attribute above), not the priority of the stream the graph is launched into
            const bool ok = ck(cudaGraphInstantiateWithFlags(&e, g, cudaGraphInstantiateFlagUseNodePriority), "cudaGraphInstantiate");
            cudaGraphDestroy(g);
            if (!ok) return false;
            graphs_[gkey_] = e;
            gexec_ = e;
        } else if (gmode_ != 2) {
            return true;
        }
        gmode_ = 0;
        ++graph_launches;
        return ck(cudaGraphLaunch(gexec_, stream_), "cudaGraphLaunch");
    }
    void graph_abort() {
        if (gmode_ == 1) {
            cudaGraph_t g = nullptr;
            cudaStreamEndCapture(stream_, &g);
            if (g) cudaGraphDestroy(g);
            cudaGetLastError();
        }
        gmode_ = 0;
        cur_ = 0;
    }
    void graph_clear() {
        if (device_ < 0) return;
        if (stream_) cudaStreamSynchronize(stream_);
        for (auto &kv : graphs_) cudaGraphExecDestroy(kv.second);
        graphs_.clear();
    }

    cudaStream_t stream() const { return stream_; }
    // 0: none; 1: a halo wait timed out; 2: an IB exchange wait did (set by the wait kernels, sticky)
    int peer_timeout() const { return timeout_word_ ? *timeout_word_ : 0; }

private:
    bool ck(cudaError_t rc, const char *what) {
        if (rc == cudaSuccess) return true;
        err = std::string(what) + ": " + cudaGetErrorString(rc);
        return false;
    }
    void *open_ipc(const unsigned char *bytes) {
        for (auto &o : opened_)
            if (std::memcmp(o.first.data(), bytes, 64) == 0) return o.second;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, bytes, 64);
        void *p = nullptr;
        if (!ck(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle")) return nullptr;
        opened_.emplace_back(std::string(reinterpret_cast<const char *>(bytes), 64), p);
        return p;
    }

    int device_ = -1;
    cudaStream_t stream_ = nullptr, copy_s_ = nullptr;
    cudaEvent_t tev_[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    int tpair_ = 0;
    std::vector<std::pair<std::string, void *>> opened_;
    static constexpr int kStreams = 6;      // 0/1 high priority (main + thin branch), 2/3 low (far branch + its thin branch),
                                            // 4/5 high (halo branch of a slab: boundary planes -> push -> signal, + its thin branch)
    cudaStream_t s_[kStreams] = {};
    cudaEvent_t fork_ev_[kStreams] = {}, join_ev_[kStreams] = {};
    int cur_ = 0;
    int prio_low_ = 0, prio_high_ = 0;
    volatile int *timeout_word_ = nullptr;     // pinned + mapped (UVA: the same pointer is valid on the device)
    unsigned long long peer_timeout_ns_ = 20ull * 1000ull * 1000ull * 1000ull;
    bool debug_sync_ = false;
    int sm_count_ = 148;
    bool coop_ = false;
    int gmode_ = 0;                 // 0 direct, 1 capturing, 2 replaying
    GraphKey gkey_{};
    cudaGraphExec_t gexec_ = nullptr;
    std::map<GraphKey, cudaGraphExec_t> graphs_;
    static constexpr int kNamedEvents = 4;
    cudaEvent_t named_[kNamedEvents] = {nullptr, nullptr, nullptr, nullptr};
    static constexpr int kMaxMarks = 8192;
    std::vector<cudaEvent_t> marks_[2];
    int nmarks_[2] = {0, 0};
};

}  // namespace fg
