// dev_host.hpp — TEST INFRASTRUCTURE: a device policy that runs the kernel bodies of
// gym-fish_b200/csrc/*.cuh in plain CPU loops (block by block, thread by thread).
// It exists so that AA-pattern indexing, bounce-back, z-face ops and the IB kernels can be checked
// against the fp64 oracle in this GPU-less container before GPU time is spent.  It is never loaded by
// the product (gym-fish_b200/_abi.py knows only the CUDA library); results from it are not benchmarks.
#pragma once
#include <atomic>
#include <chrono>
#include <deque>
#include <map>
#include <vector>
#include <functional>
#include <memory>
#include <thread>
#include <cstdio>
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
    long long graph_launches = 0;
    int sm_count() const { return 148; }
    static int interp_spread_blocks(int n_markers) { return (n_markers + 1) / 2; }
    void range_push(const char *) {}
    void range_pop() {}
    long long n_submit = 0, n_record = 0, n_wait = 0;      // FG_EMU_DEBUG: operations queued (what the CUDA policy would issue as API calls)
    int wait_ms = 300;   // how long a slab waits for its neighbour's halo before reporting it is behind

    // ---- streams.  Default (FG_EMU_SCHED unset): every operation runs at once, in submission order.
    // FG_EMU_SCHED=low | high | rand:<seed>: operations are QUEUED per stream like on the GPU (launches on the current
    // stream; copies, memsets, event records, neighbour waits / signals on stream 0) and executed only when the host
    // waits for something — in an order that respects nothing but the recorded event dependencies, preferring the lowest
    // / highest / a random runnable stream.  A missing fork / join between two launches that touch the same data then
    // shows up as a result that depends on the policy (tests/test_stream_order.py); work left on a side stream at a
    // synchronisation is reported as an error.  Nothing stays queued across ABI calls (toc_record drains).
    static constexpr int kStreams = 7;      // 0-5 as in dev_cuda.cuh, 6 = its copy stream
    static constexpr int kCopy = 6;
    bool init(int, std::string &) {
        if (const char *e = std::getenv("FG_EMU_SCHED")) {
            const std::string v(e);
            if (v == "low") sched_ = 1;
            else if (v == "high") sched_ = 2;
            else if (v.rfind("rand", 0) == 0) { sched_ = 3; rng_ = v.size() > 5 ? std::strtoull(v.c_str() + 5, nullptr, 10) * 2654435761ull + 1 : 12345; }
        }
        graphs_on_ = std::getenv("FG_EMU_GRAPHS") != nullptr;
        return true;
    }
    void enter() {}
    void shutdown() { drain_all(); }
    void *alloc(size_t bytes, std::string &e) {
        void *p = std::malloc(bytes ? bytes : 1);
        if (!p) e = "host emulation: malloc failed";
        return p;
    }
    void free(void *p) { if (p) drain_all(); std::free(p); }       // cudaFree synchronises the device
    bool h2d(void *d, const void *s, size_t n) { if (!drain_stream0()) return false; std::memcpy(d, s, n); return true; }
    bool d2h(void *d, const void *s, size_t n) { if (!drain_stream0()) return false; std::memcpy(d, s, n); return true; }
    bool zero(void *d, size_t n) { return submit(0, [=] { std::memset(d, 0, n); return true; }); }
    bool sync() {
        if (!drain_stream0()) return false;
        for (int s = 1; s < kStreams; ++s)
            if (!q_[s].empty()) { err = "host emulation: work on stream " + std::to_string(s) + " was never joined into stream 0 before a synchronisation"; return false; }
        return !failed_;
    }
    void *alloc_host(size_t bytes, std::string &e) { return alloc(bytes, e); }
    void free_host(void *p) { if (p) drain_all(); std::free(p); }
    bool h2d_async(void *d, const void *s, size_t n) { return submit(0, [=] { std::memcpy(d, s, n); return true; }); }   // reads the pinned buffer when it RUNS
    bool d2h_async(void *d, const void *s, size_t n) { return submit(0, [=] { std::memcpy(d, s, n); return true; }); }
    bool ev_record(int id) {
        if (gmode_ == 2) return true;
        if (gmode_ == 1) { graph_->ops.push_back(GOp{0, 1, nullptr, graph_->n_events++, id}); return true; }
        named_[id] = record(0);
        return true;
    }
    bool ev_sync(int id) {
        if (id < 0 || id >= 4 || !named_[id]) return true;
        std::shared_ptr<bool> e = named_[id];
        return drain_until([e] { return *e; });
    }
    // dev_cuda.cuh upload_early: the copy runs on its own stream behind event `after`, stream 0 waits for it
    bool upload_early(void *d, const void *s, size_t n, int done, int after) {
        if (gmode_ != 0) { err = "upload_early inside a graph capture"; return false; }
        if (after >= 0 && after < 4 && named_[after]) wait(kCopy, named_[after]);
        if (!submit(kCopy, [=] { std::memcpy(d, s, n); return true; })) return false;
        named_[done] = record(kCopy);
        wait(0, named_[done]);
        return true;
    }
    void tic() { t0_ = std::chrono::steady_clock::now(); }
    void marks_reset() {}
    void mark(int) {}
    double marks_elapsed(int) { return 0.0; }
    void toc_record() {
        drain_all();
        t1_ = std::chrono::steady_clock::now();
        if (std::getenv("FG_EMU_DEBUG")) std::fprintf(stderr, "[emu] call: %lld submits, %lld event records, %lld stream waits so far\n", n_submit, n_record, n_wait);
    }
    double toc_elapsed(bool, bool &ok) { ok = true; return std::chrono::duration<double, std::milli>(t1_ - t0_).count(); }

    template <class K, class P>
    bool launch(Dim3 g, const P &p) {
        ++launches;
        return submit(cur_, [=] {
            for (int bz = 0; bz < g.z; ++bz)
                for (int by = 0; by < g.y; ++by)
                    for (int bx = 0; bx < g.x; ++bx)
                        for (int tx = 0; tx < K::kThreads; ++tx) K::run(p, bx, by, bz, tx);
            return true;
        });
    }

    // block-phased kernels: CTA barrier between the phases -> per block: all threads of phase 0, then phase 1, ...
    template <class K, class P>
    bool launch_block_phased(Dim3 g, const P &p) {
        ++launches;
        return submit(cur_, [=] {
            for (int bz = 0; bz < g.z; ++bz)
                for (int by = 0; by < g.y; ++by)
                    for (int bx = 0; bx < g.x; ++bx)
                        for (int ph = 0; ph < K::kBlockPhases; ++ph)
                            for (int tx = 0; tx < K::kThreads; ++tx) K::run(p, bx, by, bz, tx, ph);
            return true;
        });
    }

    // ticket-scheduled kernels: on the GPU the phases interleave along a wavefront; here every CTA of phase 0, then phase 1
    template <class K, class P>
    bool launch_ticketed(Dim3 g, const P &p) {
        ++launches;
        return submit(cur_, [=] {
            for (int ph = 0; ph < K::kGridPhases; ++ph)
                for (int bz = 0; bz < g.z; ++bz)
                    for (int by = 0; by < g.y; ++by)
                        for (int bx = 0; bx < g.x; ++bx)
                            for (int tx = 0; tx < K::kThreads; ++tx) K::run(p, bx, by, bz, tx, ph);
            return true;
        });
    }
    bool zero_on_current(void *d, size_t n) { return submit(cur_, [=] { std::memset(d, 0, n); return true; }); }

    // phased kernels (one cooperative launch on the GPU): phases in order, every item of a phase before the next
    bool supports_phased() const { return true; }
    template <class K, class P>
    bool launch_phased(long long, const P &p) {
        ++launches;
        return submit(0, [=] {
            for (int ph = 0; ph < K::kPhases; ++ph) {
                const long long n = K::items(p, ph);
                for (long long i = 0; i < n; ++i) K::item(p, ph, i);
            }
            return true;
        });
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
        int *t[8] = {};
        for (int i = 0; i < n && i < 8; ++i) t[i] = targets ? targets[i] : nullptr;
        return submit(0, [=] {
            const int v = mine[0] + 1;
            std::atomic_thread_fence(std::memory_order_release);
            for (int i = 0; i < 8; ++i)
                if (t[i]) *static_cast<volatile int *>(t[i]) = v;
            if (bump) mine[0] = v;
            return true;
        });
    }
    bool wait_counters(int *mine, int *const *sources, int n) {
        int *src[8] = {};
        for (int i = 0; i < n && i < 8; ++i) src[i] = sources[i];
        return submit(0, [=] {
            const int v = mine[0] + 1;
            const auto t0 = std::chrono::steady_clock::now();
            for (int i = 0; i < 8; ++i) {
                if (!src[i]) continue;
                while (*static_cast<volatile int *>(src[i]) < v) {
                    if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(20)) {
                        err = "host emulation: timed out waiting for another rank's exchange";
                        fail_kind_ = 2;
                        return false;
                    }
                    std::this_thread::yield();
                }
            }
            std::atomic_thread_fence(std::memory_order_acquire);
            return true;
        });
    }
    void close_peers() {}
    int peer_timeout() { drain_all(); return fail_kind_; }   // 1: a halo wait gave up, 2: an exchange wait did (sticky, like the CUDA word)
    int current() const { return cur_; }
    // fork_to(s): stream s waits for what is queued on the current stream so far, and becomes current;
    // join_from(s): the current stream waits for what is queued on s so far (dev_cuda.cuh has the same contract)
    bool fork_to(int s) {
        const std::shared_ptr<bool> e = record(cur_);
        wait(s, e);
        cur_ = s;
        return true;
    }
    bool switch_to(int s) { cur_ = s; return true; }
    bool join_from(int s) {
        const std::shared_ptr<bool> e = record(s);
        wait(cur_, e);
        return true;
    }
    // ---- CUDA graphs (FG_EMU_GRAPHS=1; off by default: everything is launched directly).  Capture RECORDS the operations
    // with the arguments they had at capture time and runs nothing; graph_end launches the recording; for a known key the
    // host code runs again but what it submits is DROPPED and the first recording is launched instead — exactly what
    // dev_cuda.cuh does, so a launch argument that changes between substeps without being part of the key gives a wrong
    // result here, too.
    bool graph_begin(const GraphKey &key) {
        if (!graphs_on_) return false;
        auto it = graphs_.find(key);
        if (it != graphs_.end()) { gmode_ = 2; graph_ = &it->second; if (std::getenv("FG_EMU_DEBUG")) std::fprintf(stderr, "[emu] replay graph %llx\n", (unsigned long long)key.w[0]); return true; }
        if (graphs_.size() >= 64) graph_clear();
        graph_ = &graphs_[key];
        if (std::getenv("FG_EMU_DEBUG")) std::fprintf(stderr, "[emu] capture graph %llx\n", (unsigned long long)key.w[0]);
        gkey_ = key;
        gmode_ = 1;
        cur_ = 0;
        cap_ev_.clear();
        cap_keep_.clear();
        return true;
    }
    bool graph_end() {
        if (gmode_ == 0) return true;
        gmode_ = 0;
        ++graph_launches;
        std::vector<std::shared_ptr<bool>> inst(graph_->n_events);
        for (auto &e : inst) e = std::make_shared<bool>(sched_ == 0);
        for (const GOp &op : graph_->ops) {
            if (op.kind == 0) { if (!submit(op.stream, op.fn)) return false; }
            else if (op.kind == 1) {
                if (sched_ != 0) q_[op.stream].push_back(Op{1, nullptr, inst[op.ev]});
                if (op.named >= 0) named_[op.named] = inst[op.ev];
            } else wait(op.stream, inst[op.ev]);
        }
        return true;
    }
    void graph_abort() {
        if (gmode_ == 1) graphs_.erase(gkey_);
        gmode_ = 0;
        cur_ = 0;
    }
    void graph_clear() { drain_all(); graphs_.clear(); }
    bool signal_flags(int *mine, int *lo, int *hi) {
        return submit(cur_, [=] {
            const int value = ++mine[2];
            std::atomic_thread_fence(std::memory_order_release);
            if (lo) *static_cast<volatile int *>(lo) = value;
            if (hi) *static_cast<volatile int *>(hi) = value;
            return true;
        });
    }
    bool wait_flags(const int *flags, bool lo, bool hi) {
        if (!lo && !hi) return true;
        return submit(0, [=] {
            const int value = flags[2];
            const auto t0 = std::chrono::steady_clock::now();
            const volatile int *vf = flags;
            while ((lo && vf[0] < value) || (hi && vf[1] < value)) {
                if (std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(wait_ms)) {
                    err = "host emulation: neighbour slab is behind (step the slabs in lock-step or from separate threads)";
                    fail_kind_ = 1;
                    return false;
                }
                std::this_thread::yield();
            }
            std::atomic_thread_fence(std::memory_order_acquire);
            return true;
        });
    }

private:
    // ---- queued execution
    struct Op {
        int kind;                      // 0: run fn, 1: record ev, 2: wait for ev
        std::function<bool()> fn;
        std::shared_ptr<bool> ev;
    };
    bool submit(int stream, std::function<bool()> fn) {
        ++n_submit;
        if (gmode_ == 2) return true;                                             // replaying: dropped
        if (gmode_ == 1) { graph_->ops.push_back(GOp{stream, 0, std::move(fn), -1, -1}); return true; }
        if (sched_ == 0) {
            if (!fn()) { failed_ = true; return false; }
            return true;
        }
        q_[stream].push_back(Op{0, std::move(fn), nullptr});
        return true;
    }
    std::shared_ptr<bool> record(int stream) {
        ++n_record;
        auto e = std::make_shared<bool>(sched_ == 0);
        if (gmode_ == 2) return e;
        if (gmode_ == 1) {
            cap_keep_.push_back(e);                                               // keeps the address unique for the capture
            cap_ev_[e.get()] = graph_->n_events;
            graph_->ops.push_back(GOp{stream, 1, nullptr, graph_->n_events++, -1});
            return e;
        }
        if (sched_ != 0) q_[stream].push_back(Op{1, nullptr, e});
        return e;
    }
    void wait(int stream, const std::shared_ptr<bool> &e) {
        ++n_wait;
        if (gmode_ == 2) return;
        if (gmode_ == 1) {
            auto it = cap_ev_.find(e.get());
            if (it != cap_ev_.end()) graph_->ops.push_back(GOp{stream, 2, nullptr, it->second, -1});
            return;                                                                // an event from before the capture: complete
        }
        if (sched_ != 0) q_[stream].push_back(Op{2, nullptr, e});
    }
    bool runnable(int s) const { return !q_[s].empty() && (q_[s].front().kind != 2 || *q_[s].front().ev); }
    bool run_one() {
        int pick = -1;
        if (sched_ == 3) {
            int cand[kStreams], n = 0;
            for (int s = 0; s < kStreams; ++s) if (runnable(s)) cand[n++] = s;
            if (n) { rng_ = rng_ * 6364136223846793005ull + 1442695040888963407ull; pick = cand[(rng_ >> 33) % n]; }
        } else if (sched_ == 2) {
            for (int s = kStreams - 1; s >= 0 && pick < 0; --s) if (runnable(s)) pick = s;
        } else {
            for (int s = 0; s < kStreams && pick < 0; ++s) if (runnable(s)) pick = s;
        }
        if (pick < 0) return false;
        Op op = std::move(q_[pick].front());
        q_[pick].pop_front();
        if (op.kind == 0) { if (!op.fn()) failed_ = true; }
        else if (op.kind == 1) *op.ev = true;
        return true;
    }
    template <class Pred>
    bool drain_until(Pred done) {
        while (!done()) {
            if (!run_one()) {
                err = "host emulation: queued streams cannot make progress (a wait on an event that is never recorded)";
                failed_ = true;
                return false;
            }
        }
        return !failed_;
    }
    bool drain_stream0() { return sched_ == 0 ? !failed_ : drain_until([this] { return q_[0].empty(); }); }
    void drain_all() {
        if (sched_ == 0) return;
        drain_until([this] {
            for (int s = 0; s < kStreams; ++s) if (!q_[s].empty()) return false;
            return true;
        });
    }

    struct GOp { int stream, kind; std::function<bool()> fn; int ev, named; };
    struct Graph { std::vector<GOp> ops; int n_events = 0; };
    bool graphs_on_ = false;
    int gmode_ = 0;                    // 0 direct, 1 capturing, 2 replaying
    Graph *graph_ = nullptr;
    GraphKey gkey_{};
    std::map<GraphKey, Graph> graphs_;
    std::map<const bool *, int> cap_ev_;
    std::vector<std::shared_ptr<bool>> cap_keep_;
    int cur_ = 0;
    int sched_ = 0;                    // 0 immediate, 1 lowest runnable stream first, 2 highest first, 3 random
    unsigned long long rng_ = 1;
    bool failed_ = false;
    int fail_kind_ = 0;
    std::deque<Op> q_[kStreams];
    std::shared_ptr<bool> named_[4];
    std::chrono::steady_clock::time_point t0_, t1_;
};

}  // namespace fg
