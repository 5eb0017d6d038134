// hanamaru_b200.cu -- host side of the C ABI of include/hanamaru_b200.h: renderer state, the
// per-batch kernel sequence (hnm_kernels.cuh, hnm_trace.cuh), resolve, counters and timing.
// Compile with -fmad=false (parity: the reference never contracts a*b+c).
#include <dlfcn.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "hnm_device.cuh"
#include "hnm_kernels.cuh"
#include "hnm_scene.cuh"

namespace hnm {
thread_local std::string g_last_error;
}
using namespace hnm;

struct KernelTimer {
    struct Rec { const char* name; cudaEvent_t a, b; };
    std::vector<Rec> pending;
    std::vector<cudaEvent_t> pool;
    std::map<std::string, std::pair<double, uint32_t>> totals;
    std::vector<std::string> order;
    cudaEvent_t get() {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
    void collect() {
        for (auto& r : pending) {
            float ms = 0;
            cudaEventSynchronize(r.b);
            cudaEventElapsedTime(&ms, r.a, r.b);
            auto it = totals.find(r.name);
            if (it == totals.end()) { totals[r.name] = {ms, 1}; order.push_back(r.name); }
            else { it->second.first += ms; it->second.second += 1; }
            pool.push_back(r.a); pool.push_back(r.b);
        }
        pending.clear();
    }
    void reset() { collect(); totals.clear(); order.clear(); }
    ~KernelTimer() { collect(); for (auto e : pool) cudaEventDestroy(e); }
};

struct hnm_renderer {
    hnm_scene* scene = nullptr;
    RParams P;
    cudaStream_t stream = nullptr;
    std::vector<void*> allocs;
    uint32_t padded_rows = 0, max_batch = 1;
    size_t cap = 0;
    double *tmp0 = nullptr, *tmp1 = nullptr;
    uint8_t* rgb8 = nullptr;
    uint8_t* rgb8_host = nullptr;  // pinned staging of the resolved image (the caller's buffer is pageable)
    bool profiling = false, trace_stats = false, per_bounce_names = false, wid_stats = false;
    bool profile_overlap = false;
    bool isaac_tmem = true;        // generation through the TMEM pipeline (k_isaac_raygen_tm); HNM_ISAAC_TMEM=0: k_isaac_raygen
    int isaac_rounds = 0;          // HNM_ISAAC_ROUNDS=k: generation CTAs of k rounds (k x 112 paths) instead of one persistent CTA per SM
    bool confirm_tma = false;      // HNM_CONFIRM_TMA=1: the TMA-staged k_confirm (A/B)
    bool confirm_pairs = true;     // k_confirm_pairs ((ray, candidate) pairs spread over the lanes); HNM_CONFIRM_PAIRS=0: one ray per lane
    bool fast_math = false;        // hnm_set_precision(HNM_PRECISION_FAST_MATH): opt-in, statistical parity only
    bool rng_midtrace = false;     // HNM_RNG_MIDTRACE=1: the prefetch is enqueued right behind a trace launch, without waiting for it
    uint64_t launches = 0;
    KernelTimer timer;
    double* ray_buf[2][6] = {};
    double* thr_buf[2][3] = {};
    uint32_t* pid_buf[2] = {};
    // Generation sets: everything k_isaac_raygen / k_rng_overflow produce for one batch (first-bounce rays, zeroed
    // radiance, RNG tail + cursor).  Two sets, so that the set of batch i+1 is filled on `rng_stream` while batch i
    // is traced and shaded on `stream`: ISAAC seeding is shared-memory-latency bound at 3.5 warps/SM and uses no
    // registers to speak of, the trace / shade kernels use no shared memory -- they co-reside on every SM.
    struct GenSet {
        double* ray[6] = {}; double* thr[3] = {}; uint32_t* pid = nullptr;
        double* L[3] = {}; uint8_t* cursor = nullptr; uint64_t* tail = nullptr;
        uint32_t* q_ovf = nullptr; uint32_t* ovf_counter = nullptr;
        uint32_t* gen_next = nullptr;  // sliced generation: next path k_isaac_raygen_tm hands out
        cudaEvent_t ready = nullptr, released = nullptr;
        bool valid = false;        // holds the generated batch (sampling_first, batch), not consumed yet
        bool ready_recorded = false, released_recorded = false;
        uint32_t sampling_first = 0, batch = 0;
    } gen[2];
    int num_gen = 1;
    cudaStream_t rng_stream = nullptr;
    bool overlap = true;           // HNM_RNG_OVERLAP=0: generate on `stream`, in order (A/B and debugging)
    bool speculate = true;         // HNM_RNG_SPECULATE=0: no prefetch across hnm_render_passes calls
    int rng_start_bounce = 1;      // HNM_RNG_START_BOUNCE=k: the prefetch is released after bounce k of the current batch (measured best of 0..4)
    cudaEvent_t rng_gate = nullptr;
    // Sliced generation (HNM_RNG_SLICES=k, default 4; 0 = one launch released after bounce HNM_RNG_START_BOUNCE): the generation of the next batch runs as k stoppable launches of
    // k_isaac_raygen_tm, one beside the confirm / shade kernels of each of the first k bounces -- never beside k_trace,
    // which competes with it for the same issue slots and integer pipe -- and one last launch that finishes the set.
    int shade_threads = 256;         // HNM_SHADE_THREADS=128: CTAs of the shade / resolve kernels (finer grain beside a generation slice)
    int rng_slices = 4;
    uint32_t* gen_stop = nullptr;    // device word: stop level (k_gen_stop raises it, a slice of level <= it winds down)
    uint32_t gen_epoch = 0;          // level of the last stoppable slice
    cudaEvent_t slice_open[8] = {}, slice_done[8] = {};
    uint64_t gen_wasted = 0;       // speculative generations that were never consumed
    int sm_count = 148;
    CandLists cand = {};           // candidate lists of the rays in flight: camera rays [0, cap), shadow rays [cap, cap + scap)
    // NEE accepts a shadow hit iff |hit - light sample|^2 < 4 * offset (`norm()` is the squared length,
    // src/vector.rs:35-37,89-91; src/renderer.rs:282): only hits within sqrt(4 * offset) of the sample's distance
    // matter.  Slack covers that plus the f32 roundings.
    float tmax_slack = 0.0f;
    int trace_blocks_per_sm = HNM_TRACE_MIN_BLOCKS;  // persistent CTAs per SM = what the register budget allows
    cudaEvent_t marks[16] = {};
    std::function<void()> prefetch_hook;  // set by run_batch while a batch is being enqueued
    // multi-GPU gather target (hnm_group_* on device 0 of the group, hnm_dist_* on every rank): the shards of all ranks,
    // and the full image in row order
    double* gathered = nullptr;
    double* full = nullptr;
    cudaEvent_t resolve_done = nullptr;  // hnm_resolve_begin / hnm_resolve_end
    bool resolve_pending = false;
    void* nccl_comm = nullptr;     // ncclComm_t of hnm_dist_init (owned) or of the attached hnm_comm (borrowed)
    bool owns_comm = false;
    uint32_t dist_rank = 0, dist_nranks = 0;
};

struct hnm_comm {
    void* comm = nullptr;  // ncclComm_t
    int device = 0;
    uint32_t rank = 0, num_ranks = 0;
};

struct hnm_group {
    std::vector<hnm_scene*> scenes;
    std::vector<hnm_renderer*> rs;
    std::vector<cudaEvent_t> copied;  // per member: its shard has arrived on device 0
    uint32_t W = 0, H = 0;
};

namespace {

template <typename F>
void launch_timed(hnm_renderer* r, const char* name, F&& f, cudaStream_t on = nullptr) {
    r->launches++;
    if (!r->profiling) { f(); return; }
    if (!on) on = r->stream;
    cudaEvent_t a = r->timer.get(), b = r->timer.get();
    cudaEventRecord(a, on);
    f();
    cudaEventRecord(b, on);
    r->timer.pending.push_back({name, a, b});
}

template <typename T>
int dev_alloc(std::vector<void*>& allocs, T** out, size_t count) {
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(count * sizeof(T), 16));
    if (e != cudaSuccess) return set_error(HNM_ERR_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    allocs.push_back(p);
    *out = (T*)p;
    return 0;
}

// One device allocation for all wavefront buffers of a renderer (round 1 made ~80 cudaMalloc calls, 0.2-0.8 s of
// hnm_renderer_create at N = 8): the layout function runs twice, first to measure, then to hand out addresses.
struct Arena {
    char* base = nullptr;
    size_t off = 0;
    template <typename T>
    void take(T** out, size_t count) {
        const size_t bytes = (std::max<size_t>(count * sizeof(T), 16) + 255) & ~(size_t)255;
        *out = base ? (T*)(base + off) : nullptr;
        off += bytes;
    }
};

// which of the two ray queues the next launches read (`in`) and write (`out`)
void select_buffers(hnm_renderer* r, int in) {
    RParams& P = r->P;
    for (int k = 0; k < 6; k++) { P.rin[k] = r->ray_buf[in][k]; P.rout[k] = r->ray_buf[in ^ 1][k]; }
    for (int k = 0; k < 3; k++) { P.tin[k] = r->thr_buf[in][k]; P.tout[k] = r->thr_buf[in ^ 1][k]; }
    P.pin = r->pid_buf[in]; P.pout = r->pid_buf[in ^ 1];
}
// first bounce: read the generated rays of set `g`, write ray queue 0
void select_first_bounce(hnm_renderer* r, const hnm_renderer::GenSet& g) {
    RParams& P = r->P;
    for (int k = 0; k < 6; k++) { P.rin[k] = g.ray[k]; P.rout[k] = r->ray_buf[0][k]; }
    for (int k = 0; k < 3; k++) { P.tin[k] = g.thr[k]; P.tout[k] = r->thr_buf[0][k]; }
    P.pin = g.pid; P.pout = r->pid_buf[0];
}
// the per-path state of the batch lives in its generation set
void bind_gen_set(RParams& P, const hnm_renderer::GenSet& g) {
    for (int k = 0; k < 3; k++) P.L[k] = g.L[k];
    P.cursor = g.cursor; P.tail = g.tail; P.q_ovf = g.q_ovf; P.ovf_counter = g.ovf_counter;
}

// ISAAC seeding + lens sampling + first-bounce rays of batch (sampling_first, batch) into set `g`, on stream `on`:
// generate_begin, any number of generate_slice launches (each carries on where the previous one stopped), generate_finish.
RParams gen_params(hnm_renderer* r, hnm_renderer::GenSet& g, uint32_t sampling_first, uint32_t batch) {
    RParams G = r->P;
    G.batch = batch;
    G.sampling_first = sampling_first;
    G.N = batch * G.npix * G.spp;
    bind_gen_set(G, g);
    for (int k = 0; k < 6; k++) G.rout[k] = g.ray[k];
    for (int k = 0; k < 3; k++) G.tout[k] = g.thr[k];
    G.pout = g.pid;
    G.gen_next = g.gen_next;
    G.gen_stop = r->gen_stop;
    G.gen_epoch = 0xFFFFFFFFu;
    return G;
}
int generate_begin(hnm_renderer* r, hnm_renderer::GenSet& g, cudaStream_t on) {
    HNM_CUDA(cudaMemsetAsync(g.ovf_counter, 0, sizeof(uint32_t), on));
    HNM_CUDA(cudaMemsetAsync(g.gen_next, 0, sizeof(uint32_t), on));
    (void)r;
    return 0;
}
// one launch of the generation kernel; epoch = 0xFFFFFFFF: runs until the set is complete
void generate_slice(hnm_renderer* r, RParams& G, uint32_t epoch, cudaStream_t on) {
    const size_t smem = (size_t)ISAAC_PATHS * 256 * sizeof(uint64_t);
    G.gen_epoch = epoch;
    launch_timed(r, "isaac_raygen", [&] { k_isaac_raygen_tm<<<r->sm_count, ISAAC_TM_THREADS, smem, on>>>(G); }, on);
}
void generate_finish(hnm_renderer* r, hnm_renderer::GenSet& g, RParams& G, uint32_t sampling_first, uint32_t batch, cudaStream_t on) {
    launch_timed(r, "rng_overflow", [&] { k_rng_overflow<<<r->sm_count, 64, 0, on>>>(G); }, on);
    g.valid = true;
    g.sampling_first = sampling_first;
    g.batch = batch;
}
int generate(hnm_renderer* r, hnm_renderer::GenSet& g, uint32_t sampling_first, uint32_t batch, cudaStream_t on) {
    RParams G = gen_params(r, g, sampling_first, batch);
    int rc = generate_begin(r, g, on);
    if (rc) return rc;
    if (r->isaac_tmem) {
        generate_slice(r, G, 0xFFFFFFFFu, on);
    } else {
        const size_t smem = (size_t)ISAAC_PATHS * 256 * sizeof(uint64_t);
        // Grid: one persistent CTA per SM, or (isaac_rounds > 0) short-lived CTAs of `isaac_rounds` x 112 paths (A/B: the SM's
        // issue arbiter serves the oldest resident warps first, tools/microbench/prio.cu; measured slower)
        int igrid = r->sm_count;
        if (r->isaac_rounds > 0 && on == r->rng_stream)
            igrid = (int)std::max<uint64_t>(r->sm_count, ((uint64_t)G.N + (uint64_t)ISAAC_PATHS * r->isaac_rounds - 1) / ((uint64_t)ISAAC_PATHS * r->isaac_rounds));
        launch_timed(r, "isaac_raygen", [&] { k_isaac_raygen<<<igrid, ISAAC_THREADS, smem, on>>>(G); }, on);
    }
    generate_finish(r, g, G, sampling_first, batch, on);
    return 0;
}

TraceJob camera_job(const RParams& P, int bounce, bool classify) {
    TraceJob j;
    memset(&j, 0, sizeof(j));
    for (int k = 0; k < 6; k++) j.ray[k] = P.rin[k];
    j.hit_t = P.hit_t; j.hit_u = P.hit_u; j.hit_v = P.hit_v; j.hit_id = P.hit_id;
    j.count = &P.counters[bounce * C_STRIDE + C_RAY];
    j.slot0 = 0;
    if (classify) {
        j.cnt_miss = &P.counters[bounce * C_STRIDE + C_MISS]; j.cnt_delta = &P.counters[bounce * C_STRIDE + C_DELTA];
        j.cnt_nee = &P.counters[bounce * C_STRIDE + C_NEE];
        j.q_miss = P.q_miss; j.q_delta = P.q_delta; j.q_nee = P.q_nee;
    }
    return j;
}
TraceJob shadow_job(const RParams& P, int bounce) {
    TraceJob j;
    memset(&j, 0, sizeof(j));
    for (int k = 0; k < 6; k++) j.ray[k] = P.sray[k];
    // no hit record: k_nee_resolve confirms these rays itself
    j.count = &P.counters[bounce * C_STRIDE + C_SHADOW];
    j.tmax = P.s_tmax;
    j.slot0 = P.cap;  // the shadow rays' candidate lists follow the camera rays'
    return j;
}

// diagnostics: HNM_TRACE_NAMES=1 times every bounce's trace launch under its own name
const char* trace_name(hnm_renderer* r, int bounce) {
    static const char* names[] = {"trace_b00", "trace_b01", "trace_b02", "trace_b03", "trace_b04", "trace_b05", "trace_b06", "trace_b07",
                                  "trace_b08", "trace_b09", "trace_b10", "trace_b11", "trace_b12"};
    return (r->per_bounce_names && bounce >= 0 && bounce <= 12) ? names[bounce] : "trace";
}

void launch_nee_resolve(hnm_renderer* r, int bounce) {
    const int grid = r->sm_count * 4;
    RParams& P = r->P;
    cudaStream_t st = r->stream;
    CandLists cand = r->cand;
    const int T = r->shade_threads, G = grid * (256 / T);
    if (r->trace_stats) launch_timed(r, "nee_resolve", [&] { k_nee_resolve<true, false><<<G, T, 0, st>>>(P, cand, bounce); });
    else if (r->fast_math) launch_timed(r, "nee_resolve", [&] { k_nee_resolve<false, true><<<G, T, 0, st>>>(P, cand, bounce); });
    else launch_timed(r, "nee_resolve", [&] { k_nee_resolve<false, false><<<G, T, 0, st>>>(P, cand, bounce); });
}

// k_trace over one or two ray lists; k_confirm for the FIRST list if `confirm_first` (a shadow-ray list is confirmed by
// its consumer, k_nee_resolve)
void launch_trace(hnm_renderer* r, const char* name, const TraceJob* j0, const TraceJob* j1, uint32_t* work, int stat_segments,
                  bool confirm_first = true, bool prefetch_behind_trace = false) {
    TraceArgs A;
    memset(&A, 0, sizeof(A));
    A.work_confirm = work + (C_W_CONFIRM - C_WORK);  // same bounce row of counters[]
    A.job[0] = *j0;
    A.njobs = 1;
    if (j1) { A.job[1] = *j1; A.njobs = 2; }
    A.work = work;
    A.stats = r->P.stats;
    A.stat_segments = stat_segments;
    A.stat_nodes = S_NODES; A.stat_prims = S_PRIMS;
    A.tmax_slack = r->tmax_slack;
    A.dbg = r->P.dbg;
    A.cand = r->cand;
    const int grid = r->sm_count * r->trace_blocks_per_sm;
    const int cgrid = r->sm_count * 8;
    DScene sc = r->P.sc;
    cudaStream_t st = r->stream;
    TraceArgs C = A;
    C.njobs = 1;
    if (r->trace_stats) {
        launch_timed(r, name, [&] { k_trace<true><<<grid, TRACE_THREADS, 0, st>>>(sc, A); });
        if (confirm_first && r->confirm_pairs) launch_timed(r, "confirm", [&] { k_confirm_pairs<true><<<cgrid, 256, 0, st>>>(sc, C); });
        else if (confirm_first) launch_timed(r, "confirm", [&] { k_confirm<true><<<cgrid, 256, 0, st>>>(sc, C); });
    } else {
        launch_timed(r, name, [&] { k_trace<false><<<grid, TRACE_THREADS, 0, st>>>(sc, A); });
        if (prefetch_behind_trace && r->prefetch_hook) r->prefetch_hook();
        if (confirm_first && r->confirm_tma) launch_timed(r, "confirm", [&] { k_confirm_tma<false><<<r->sm_count * HNM_CONFIRM_MIN_BLOCKS, 256, 0, st>>>(sc, C); });
        else if (confirm_first && r->confirm_pairs) launch_timed(r, "confirm", [&] { k_confirm_pairs<false><<<cgrid, 256, 0, st>>>(sc, C); });
        else if (confirm_first) launch_timed(r, "confirm", [&] { k_confirm<false><<<cgrid, 256, 0, st>>>(sc, C); });
    }
}

// One batch of passes.  (next_first, next_batch) is the batch the caller expects to run after this one (0 = none):
// its generation is enqueued on the RNG stream BEFORE this batch's kernels, so the two overlap on the device.
int run_batch(hnm_renderer* r, uint32_t sampling_first, uint32_t batch, uint32_t next_first, uint32_t next_batch) {
    RParams& P = r->P;
    P.batch = batch;
    P.sampling_first = sampling_first;
    P.N = batch * P.npix * P.spp;
    if (P.N == 0) return 0;
    cudaStream_t st = r->stream;
    HNM_CUDA(cudaMemsetAsync(P.counters, 0, sizeof(uint32_t) * NUM_COUNTERS, st));
    const int grid = r->sm_count * 4;
    const int last = (int)P.sc.bounce_limit - 1;
    if (P.mode == HNM_MODE_PATHTRACING) {
        // per-kernel timing normally serialises the two streams (clean numbers per kernel); HNM_PROFILE_OVERLAP=1 keeps the
        // overlap on, so that the event pairs show how long each kernel takes WHILE the generation kernel is co-resident
        const bool overlap = r->overlap && (!r->profiling || r->profile_overlap) && r->num_gen == 2;
        // ---- the generation set of this batch: prefetched, or generated now
        int s = -1;
        for (int k = 0; k < r->num_gen; k++)
            if (r->gen[k].valid && r->gen[k].sampling_first == sampling_first && r->gen[k].batch == batch) s = k;
        if (s < 0) {
            s = 0;
            for (int k = 0; k < r->num_gen; k++) if (!r->gen[k].valid) s = k;
            hnm_renderer::GenSet& g = r->gen[s];
            if (g.valid) r->gen_wasted++;
            // the set's last consumer ran on `st`, its last producer possibly on the RNG stream
            if (g.ready_recorded) HNM_CUDA(cudaStreamWaitEvent(st, g.ready, 0));
            int rc = generate(r, g, sampling_first, batch, st);
            if (rc) return rc;
        } else if (r->gen[s].ready_recorded) {
            HNM_CUDA(cudaStreamWaitEvent(st, r->gen[s].ready, 0));
        }
        hnm_renderer::GenSet& g = r->gen[s];
        g.valid = false;  // consumed by this batch
        // ---- prefetch the next batch into the other set (waits until that set's last consumer is done)
        auto prefetch_next = [&]() -> int {
            hnm_renderer::GenSet& o = r->gen[s ^ 1];
            if (!(o.valid && o.sampling_first == next_first && o.batch == next_batch)) {
                if (o.valid) r->gen_wasted++;
                if (o.released_recorded) HNM_CUDA(cudaStreamWaitEvent(r->rng_stream, o.released, 0));
                int rc = generate(r, o, next_first, next_batch, r->rng_stream);
                if (rc) return rc;
                HNM_CUDA(cudaEventRecord(o.ready, r->rng_stream));
                o.ready_recorded = true;
            }
            return 0;
        };
        const bool want_prefetch = overlap && next_batch > 0;
        int hook_rc = 0;
        // ---- sliced prefetch: stoppable generation launches beside the confirm / shade kernels of the first bounces
        hnm_renderer::GenSet& o = r->gen[s ^ 1];
        // (the stop level is a 32-bit counter that only grows; 0xFFFFFFFF means "never stop": slicing simply ends long before)
        const bool sliced = want_prefetch && r->isaac_tmem && r->rng_slices > 0 && r->gen_epoch < 0xFFFFFF00u &&
                            !(o.valid && o.sampling_first == next_first && o.batch == next_batch);
        RParams G;
        uint32_t slice_epoch = 0;
        int slices_left = sliced ? std::min(std::min(r->rng_slices, 8), last) : 0;
        bool gen_finished = false;
        auto finish_generation = [&]() {
            generate_slice(r, G, 0xFFFFFFFFu, r->rng_stream);  // whatever is left, not stoppable
            generate_finish(r, o, G, next_first, next_batch, r->rng_stream);
            if (cudaEventRecord(o.ready, r->rng_stream) != cudaSuccess) hook_rc = set_error(HNM_ERR_CUDA, "cudaEventRecord failed");
            o.ready_recorded = true;
            gen_finished = true;
        };
        if (sliced) {
            if (o.valid) r->gen_wasted++;
            o.valid = false;
            if (o.released_recorded) HNM_CUDA(cudaStreamWaitEvent(r->rng_stream, o.released, 0));
            G = gen_params(r, o, next_first, next_batch);
            int rc = generate_begin(r, o, r->rng_stream);
            if (rc) return rc;
            r->prefetch_hook = [&] {
                // called right behind this bounce's k_trace launch: the slice starts when the trace has finished ...
                const int i = slices_left;
                if (cudaEventRecord(r->slice_open[i & 7], st) != cudaSuccess || cudaStreamWaitEvent(r->rng_stream, r->slice_open[i & 7], 0) != cudaSuccess) {
                    hook_rc = set_error(HNM_ERR_CUDA, "slice events failed");
                    return;
                }
                slice_epoch = ++r->gen_epoch;
                generate_slice(r, G, slice_epoch, r->rng_stream);
                if (cudaEventRecord(r->slice_done[i & 7], r->rng_stream) != cudaSuccess) hook_rc = set_error(HNM_ERR_CUDA, "slice events failed");
            };
        } else {
            r->prefetch_hook = [&] { hook_rc = prefetch_next(); };
            if (want_prefetch && r->rng_start_bounce <= 0) { int rc = prefetch_next(); if (rc) return rc; }
        }
        bind_gen_set(P, g);
        const int ST = r->shade_threads, SG = grid * (256 / ST);
        // k_shade_surf: its CTA is the domain of the octant grouping
        const int SUT = (HNM_SURF_THREADS > 256) ? HNM_SURF_THREADS : ST, SUG = (HNM_SURF_THREADS > 256) ? grid * 256 / HNM_SURF_THREADS : SG;
        launch_timed(r, "batch_begin", [&] { k_batch_begin<<<1, 1, 0, st>>>(P); });
        for (int b = 1; b <= last; b++) {
            if (b == 1) select_first_bounce(r, g);
            else select_buffers(r, b & 1);  // bounce 2 reads queue 0 (written by bounce 1), bounce 3 queue 1, ...
            TraceJob cam = camera_job(P, b, true);
            // HNM_RNG_MIDTRACE: the generation kernel is enqueued right BEHIND this bounce's trace launch with no dependency
            // on it, so that its CTAs are placed while the trace CTAs are already resident (issue arbitration)
            const bool slice_here = sliced && slices_left > 0;
            const bool mid = slice_here || (!sliced && want_prefetch && r->rng_midtrace && b == r->rng_start_bounce);
            if (b > 1) {
                TraceJob sh = shadow_job(P, b - 1);
                launch_trace(r, trace_name(r, b), &cam, &sh, &P.counters[b * C_STRIDE + C_WORK], S_SEGMENTS, true, mid);
                launch_nee_resolve(r, b - 1);
            } else {
                launch_trace(r, trace_name(r, b), &cam, nullptr, &P.counters[b * C_STRIDE + C_WORK], S_SEGMENTS, true, mid);
            }
            if (r->fast_math) {
                launch_timed(r, "shade_miss", [&] { k_shade_miss<true><<<SG, ST, 0, st>>>(P, b); });
                launch_timed(r, "shade_delta", [&] { k_shade_surf<false, true><<<SUG, SUT, 0, st>>>(P, b); });
                launch_timed(r, "shade_nee", [&] { k_shade_surf<true, true><<<SUG, SUT, 0, st>>>(P, b); });
            } else {
                launch_timed(r, "shade_miss", [&] { k_shade_miss<false><<<SG, ST, 0, st>>>(P, b); });
                launch_timed(r, "shade_delta", [&] { k_shade_surf<false, false><<<SUG, SUT, 0, st>>>(P, b); });
                launch_timed(r, "shade_nee", [&] { k_shade_surf<true, false><<<SUG, SUT, 0, st>>>(P, b); });
            }
            if (slice_here) {
                // ... and winds down when this bounce's shade kernels are done; the next trace waits for it
                const int i = slices_left;
                k_gen_stop<<<1, 1, 0, st>>>(r->gen_stop, slice_epoch);
                HNM_CUDA(cudaStreamWaitEvent(st, r->slice_done[i & 7], 0));
                slices_left--;
                if (slices_left == 0) finish_generation();  // the rest runs beside the thin late bounces
            }
            if (!sliced && want_prefetch && !r->rng_midtrace && b == r->rng_start_bounce) {
                // the generation of the next batch starts here: the thin late bounces leave the SMs under-used
                HNM_CUDA(cudaEventRecord(r->rng_gate, st));
                HNM_CUDA(cudaStreamWaitEvent(r->rng_stream, r->rng_gate, 0));
                int rc = prefetch_next();
                if (rc) return rc;
            }
        }
        if (sliced && !gen_finished) finish_generation();
        TraceJob sh = shadow_job(P, last);
        launch_trace(r, trace_name(r, last + 1), &sh, nullptr, &P.counters[(last + 1) * C_STRIDE + C_WORK], -1, false);
        launch_nee_resolve(r, last);
        launch_timed(r, "accumulate", [&] { k_accumulate<<<grid, 256, 0, st>>>(P); });
        r->prefetch_hook = nullptr;
        if (hook_rc) return hook_rc;
        HNM_CUDA(cudaEventRecord(g.released, st));
        g.released_recorded = true;
    } else {
        bind_gen_set(P, r->gen[0]);
        select_buffers(r, 1);
        launch_timed(r, "raygen_debug", [&] { k_raygen_debug<<<grid, 256, 0, st>>>(P); });
        select_buffers(r, 0);
        TraceJob cam = camera_job(P, 1, false);
        launch_trace(r, "trace", &cam, nullptr, &P.counters[1 * C_STRIDE + C_WORK], S_SEGMENTS);
        launch_timed(r, "debug_shade", [&] { k_debug_shade<<<grid, 256, 0, st>>>(P); });
        launch_timed(r, "accumulate", [&] { k_accumulate<<<grid, 256, 0, st>>>(P); });
    }
    HNM_CUDA(cudaGetLastError());
    return 0;
}

// scratch device buffers of the batch entry points: released on every return path
struct TmpDev {
    std::vector<void*> v;
    ~TmpDev() { for (auto p : v) cudaFree(p); }
    template <typename T>
    int alloc(T** out, size_t bytes) {
        void* p = nullptr;
        HNM_CUDA(cudaMalloc(&p, std::max<size_t>(bytes, 16)));
        v.push_back(p);
        *out = (T*)p;
        return 0;
    }
    template <typename T>
    int upload(T** out, const void* host, size_t bytes) {
        int rc = alloc(out, bytes);
        if (rc) return rc;
        HNM_CUDA(cudaMemcpy(*out, host, bytes, cudaMemcpyHostToDevice));
        return 0;
    }
};
static int device_sm_count(int device, int* n) {
    HNM_CUDA(cudaDeviceGetAttribute(n, cudaDevAttrMultiProcessorCount, device));
    return 0;
}


// the gather target of a sharded renderer, allocated on first use
int ensure_gather_buffers(hnm_renderer* r) {
    if (r->gathered && r->full) return 0;
    const size_t shard = (size_t)r->padded_rows * r->P.W * 3;
    double *g = nullptr, *f = nullptr;
    cudaError_t e = cudaMalloc(&g, shard * r->P.nranks * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&f, (size_t)r->P.W * r->P.H * 3 * sizeof(double));
    if (e != cudaSuccess) { cudaFree(g); cudaFree(f); return set_error(HNM_ERR_NOMEM, std::string("gather buffers: ") + cudaGetErrorString(e)); }
    r->allocs.push_back(g); r->allocs.push_back(f);
    r->gathered = g; r->full = f;
    return 0;
}

// ---- NCCL, bound at run time (dlopen): a single-GPU user needs no libnccl, and inside a process that already
// loaded one (torch bundles its own) the same instance is used.  Only what the one exchange step needs.
struct NcclApi {
    typedef struct { char internal[128]; } UniqueId;  // ncclUniqueId
    int (*GetUniqueId)(UniqueId*) = nullptr;
    int (*CommInitRank)(void**, int, UniqueId, int) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
    std::string why;
};
NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        // Order: the file HNM_NCCL_LIB names (the Python mirror points it at the wheel torch itself loads, so that both
        // agree whichever comes first: two different libnccl.so.2 in one process clash by soname); an instance the process
        // has already loaded; the system's.
        void* h = nullptr;
        if (const char* e = getenv("HNM_NCCL_LIB")) { if (*e) h = dlopen(e, RTLD_NOW | RTLD_GLOBAL); }
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        if (!h) for (const char* name : {"libnccl.so.2", "libnccl.so"}) { h = dlopen(name, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
        if (!h) { api.why = "libnccl.so.2 not found"; return; }
        api.GetUniqueId = (int (*)(NcclApi::UniqueId*))dlsym(h, "ncclGetUniqueId");
        api.CommInitRank = (int (*)(void**, int, NcclApi::UniqueId, int))dlsym(h, "ncclCommInitRank");
        api.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(h, "ncclAllGather");
        api.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
        api.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
        api.ok = api.GetUniqueId && api.CommInitRank && api.AllGather && api.CommDestroy && api.GetErrorString;
        if (!api.ok) api.why = "libnccl lacks an expected symbol";
    });
    return api;
}
constexpr int NCCL_FLOAT64 = 8;  // ncclFloat64 / ncclDouble (nccl.h: ncclDataType_t)
#define HNM_NCCL(call)                                                                                     \
    do {                                                                                                   \
        int e__ = (call);                                                                                  \
        if (e__ != 0) return set_error(HNM_ERR_CUDA, std::string(#call) + ": " + nccl().GetErrorString(e__)); \
    } while (0)
}  // namespace

extern "C" {

const char* hnm_last_error(void) { return g_last_error.c_str(); }
uint32_t hnm_abi_version(void) { return HNM_ABI_VERSION; }
int hnm_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return set_error(HNM_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    return n;
}

int hnm_scene_create(const hnm_scene_desc* desc, int device, hnm_scene** out) { return scene_create(desc, device, out); }
void hnm_scene_destroy(hnm_scene* s) { scene_free(s); }

void hnm_renderer_destroy(hnm_renderer* r) {
    if (!r) return;
    cudaSetDevice(r->scene->device);
    if (r->rng_stream) cudaStreamSynchronize(r->rng_stream);
    if (r->stream) cudaStreamSynchronize(r->stream);
    r->timer.collect();
    if (r->nccl_comm && r->owns_comm && nccl().ok) nccl().CommDestroy(r->nccl_comm);
    for (auto p : r->allocs) cudaFree(p);
    if (r->rgb8_host) cudaFreeHost(r->rgb8_host);
    for (auto e : r->marks) if (e) cudaEventDestroy(e);
    for (auto& g : r->gen) {
        if (g.ready) cudaEventDestroy(g.ready);
        if (g.released) cudaEventDestroy(g.released);
    }
    if (r->rng_gate) cudaEventDestroy(r->rng_gate);
    if (r->resolve_done) cudaEventDestroy(r->resolve_done);
    for (int k = 0; k < 8; k++) { if (r->slice_open[k]) cudaEventDestroy(r->slice_open[k]); if (r->slice_done[k]) cudaEventDestroy(r->slice_done[k]); }
    if (r->rng_stream) cudaStreamDestroy(r->rng_stream);
    if (r->stream) cudaStreamDestroy(r->stream);
    delete r;
}

int hnm_renderer_create(hnm_scene* scene, const hnm_camera* camera, uint32_t width, uint32_t height, int mode, const hnm_shard* shard,
                        uint32_t max_batch, hnm_renderer** out) {
    if (!scene || !camera || !out) return set_error(HNM_ERR_INVALID, "null argument");
    *out = nullptr;
    if (width == 0 || height == 0 || (uint64_t)width * height > (1ull << 28)) return set_error(HNM_ERR_INVALID, "bad resolution");
    if (mode < HNM_MODE_PATHTRACING || mode > HNM_MODE_DEBUG_FOCALPLANE) return set_error(HNM_ERR_INVALID, "bad mode");
    hnm_shard sh = {0, 1, 8, 0};
    if (shard) sh = *shard;
    if (sh.num_ranks == 0 || sh.rank >= sh.num_ranks || sh.tile_rows == 0) return set_error(HNM_ERR_INVALID, "bad shard");
    if (sh.num_ranks == 1) sh.tile_rows = height;  // one tile: local row == image row
    HNM_CUDA(cudaSetDevice(scene->device));
    int sm_count = 0;
    HNM_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, scene->device));  // before `new`: nothing to release on failure
    hnm_renderer* r = new hnm_renderer();
    r->scene = scene;
    r->sm_count = sm_count;
    RParams& P = r->P;
    memset(&P, 0, sizeof(P));
    P.sc = scene->d;
    P.cam = *camera;
    P.W = width; P.H = height;
    P.ss = scene->config.supersampling; P.spp = P.ss * P.ss;
    P.rank = sh.rank; P.nranks = sh.num_ranks; P.tile_rows = sh.tile_rows;
    P.mode = mode;
    P.tail_k = RNG_TAIL;
    if (const char* e = getenv("HNM_RNG_TAIL_K")) { int k = atoi(e); if (k >= 2 && k <= RNG_TAIL) P.tail_k = k & ~1; }
    if (const char* e = getenv("HNM_TRACE_STATS")) r->trace_stats = atoi(e) != 0;
    if (const char* e = getenv("HNM_TRACE_NAMES")) r->per_bounce_names = atoi(e) != 0;
    if (const char* e = getenv("HNM_WID_STATS")) r->wid_stats = atoi(e) != 0;
    if (const char* e = getenv("HNM_RNG_MIDTRACE")) r->rng_midtrace = atoi(e) != 0;
    if (const char* e = getenv("HNM_PROFILE_OVERLAP")) r->profile_overlap = atoi(e) != 0;
    if (const char* e = getenv("HNM_CONFIRM_TMA")) r->confirm_tma = atoi(e) != 0;
    if (const char* e = getenv("HNM_CONFIRM_PAIRS")) r->confirm_pairs = atoi(e) != 0;
    if (const char* e = getenv("HNM_ISAAC_TMEM")) r->isaac_tmem = atoi(e) != 0;
    if (const char* e = getenv("HNM_SHADE_THREADS")) { int t = atoi(e); if (t == 64 || t == 128 || t == 256) r->shade_threads = t; }
    if (const char* e = getenv("HNM_RNG_SLICES")) r->rng_slices = std::max(0, std::min(8, atoi(e)));
    if (const char* e = getenv("HNM_ISAAC_ROUNDS")) r->isaac_rounds = std::max(0, atoi(e));
    if (const char* e = getenv("HNM_TRACE_BLOCKS")) { int k = atoi(e); if (k >= 1 && k <= 16) r->trace_blocks_per_sm = k; }
    if (const char* e = getenv("HNM_RNG_OVERLAP")) r->overlap = atoi(e) != 0;
    if (const char* e = getenv("HNM_RNG_SPECULATE")) r->speculate = atoi(e) != 0;
    if (const char* e = getenv("HNM_RNG_START_BOUNCE")) r->rng_start_bounce = atoi(e);
    uint32_t ntiles = (height + sh.tile_rows - 1) / sh.tile_rows;
    uint32_t tiles_per_rank = (ntiles + sh.num_ranks - 1) / sh.num_ranks;
    r->padded_rows = tiles_per_rank * sh.tile_rows;  // equal on every rank (all-gather)
    uint32_t real_rows = 0;
    for (uint32_t lr = 0; lr < r->padded_rows; lr++)
        if (local_to_global_row_h(lr, sh.rank, sh.num_ranks, sh.tile_rows) < height) real_rows++;
    // rows that exist are a prefix of the local rows (tiles are assigned in increasing order)
    P.real_rows = real_rows;
    P.npix = real_rows * width;
    // NEE accepts a shadow hit iff (hit - sample).norm() < 4 * OFFSET, and `norm` is the SQUARED length
    // (src/vector.rs:35-37,89-91): the hit must lie within sqrt(4 * OFFSET) = 0.02 of the sample.
    r->tmax_slack = (float)(std::sqrt(4.0 * scene->config.offset) * 1.05 + 1e-5 * (double)scene->d.scene_r + 1e-6);
    if (const char* e = getenv("HNM_SHADOW_BOUNDED")) { if (atoi(e) == 0) r->tmax_slack = 3.0e38f; }  // A/B: unbounded closest hit
    size_t per_pass = (size_t)P.npix * P.spp;
    if (mode != HNM_MODE_PATHTRACING) max_batch = 1;
    if (max_batch == 0) {
        // enough paths in flight to fill the machine through the thin late bounces, bounded memory
        size_t target = 16u << 20;  // measured: 16 M paths per batch run 28 % faster per path than 8 M (thin late bounces)
        max_batch = (uint32_t)std::min<size_t>(64, std::max<size_t>(1, (target + per_pass - 1) / std::max<size_t>(per_pass, 1)));
    }
    while (max_batch > 1 && per_pass * max_batch > (1ull << 31) - 64) max_batch--;
    if (per_pass * max_batch > (1ull << 31) - 64) { delete r; return set_error(HNM_ERR_INVALID, "image too large for one batch"); }
    r->max_batch = max_batch;
    r->cap = std::max<size_t>(per_pass * max_batch, 32);
    P.cap = (uint32_t)r->cap;
    int rc = 0;
    auto bail = [&](int code) { hnm_renderer_destroy(r); return code; };
    cudaError_t ce = cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking);
    if (ce != cudaSuccess) { delete r; return set_error(HNM_ERR_CUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(ce)); }
    r->num_gen = (mode == HNM_MODE_PATHTRACING && r->overlap) ? 2 : 1;
    if (r->num_gen == 2) {
        // the generation kernel needs a whole SM's shared memory: give its CTAs priority whenever an SM drains
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        ce = cudaStreamCreateWithPriority(&r->rng_stream, cudaStreamNonBlocking, prio_hi);
        if (ce != cudaSuccess) { set_error(HNM_ERR_CUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(ce)); return bail(HNM_ERR_CUDA); }
    }
    size_t cap = r->cap;
    const size_t scap = cap * std::max<uint32_t>(scene->num_emissions, 1);
    const size_t nl = cap + (mode == HNM_MODE_PATHTRACING ? scap : 0);
    if (nl >= (1ull << 32)) { set_error(HNM_ERR_INVALID, "too many rays in flight for 32-bit list slots"); return bail(HNM_ERR_INVALID); }
    r->cand.stride = (uint32_t)nl;
    const size_t accum_n = (size_t)r->padded_rows * width * 3;
    auto layout = [&](Arena& A) {
        for (int b = 0; b < 2; b++) {
            for (int k = 0; k < 6; k++) A.take(&r->ray_buf[b][k], cap);
            for (int k = 0; k < 3; k++) A.take(&r->thr_buf[b][k], cap);
            A.take(&r->pid_buf[b], cap);
        }
        for (int s = 0; s < r->num_gen; s++) {
            hnm_renderer::GenSet& g = r->gen[s];
            for (int k = 0; k < 3; k++) A.take(&g.L[k], cap);
            A.take(&g.cursor, cap);
            if (mode == HNM_MODE_PATHTRACING) {
                for (int k = 0; k < 6; k++) A.take(&g.ray[k], cap);
                for (int k = 0; k < 3; k++) A.take(&g.thr[k], cap);
                A.take(&g.pid, cap);
                A.take(&g.tail, cap * RNG_TAIL);
                A.take(&g.q_ovf, cap);
                A.take(&g.ovf_counter, (size_t)4);
                A.take(&g.gen_next, (size_t)4);
            }
        }
        A.take(&P.hit_t, cap); A.take(&P.hit_u, cap); A.take(&P.hit_v, cap); A.take(&P.hit_id, cap);
        if (mode == HNM_MODE_PATHTRACING) {
            A.take(&P.q_miss, cap); A.take(&P.q_delta, cap); A.take(&P.q_nee, cap);
            for (int k = 0; k < 3; k++) {
                A.take(&P.ev_thr[k], cap); A.take(&P.ev_albedo[k], cap); A.take(&P.ev_emission[k], cap);
                A.take(&P.s_pos[k], scap);
            }
            A.take(&P.ev_pid, cap);
            for (int k = 0; k < 6; k++) A.take(&P.sray[k], scap);
            A.take(&P.s_bsdf, scap); A.take(&P.s_g, scap); A.take(&P.s_tmax, scap);
        }
        A.take(&r->cand.id, nl * TRACE_CAND); A.take(&r->cand.lo, nl * TRACE_CAND);
        A.take(&r->cand.n, nl); A.take(&r->cand.ub, nl);
        A.take(&P.counters, (size_t)NUM_COUNTERS);
        A.take(&r->gen_stop, (size_t)4);
        A.take(&P.stats, (size_t)S_COUNT);
        A.take(&P.accum, accum_n);
        if (r->wid_stats) A.take(&P.dbg, (size_t)8);
    };
    {
        Arena measure;
        layout(measure);
        char* base = nullptr;
        if ((rc = dev_alloc(r->allocs, &base, measure.off))) return bail(rc);
        Arena place;
        place.base = base;
        layout(place);
    }
    for (int s = 0; s < r->num_gen; s++) {
        hnm_renderer::GenSet& g = r->gen[s];
        if (cudaEventCreateWithFlags(&g.ready, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&g.released, cudaEventDisableTiming) != cudaSuccess) {
            set_error(HNM_ERR_CUDA, "cudaEventCreate failed");
            return bail(HNM_ERR_CUDA);
        }
    }
    if (cudaEventCreateWithFlags(&r->rng_gate, cudaEventDisableTiming) != cudaSuccess) { set_error(HNM_ERR_CUDA, "cudaEventCreate failed"); return bail(HNM_ERR_CUDA); }
    for (int k = 0; k < 8; k++)
        if (cudaEventCreateWithFlags(&r->slice_open[k], cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&r->slice_done[k], cudaEventDisableTiming) != cudaSuccess) {
            set_error(HNM_ERR_CUDA, "cudaEventCreate failed");
            return bail(HNM_ERR_CUDA);
        }
    bind_gen_set(P, r->gen[0]);
    ce = cudaMemsetAsync(P.accum, 0, accum_n * sizeof(double), r->stream);
    if (ce == cudaSuccess) ce = cudaMemsetAsync(P.stats, 0, S_COUNT * sizeof(unsigned long long), r->stream);
    if (ce == cudaSuccess) ce = cudaMemsetAsync(r->gen_stop, 0, sizeof(uint32_t), r->stream);
    if (ce == cudaSuccess && P.dbg) ce = cudaMemsetAsync(P.dbg, 0, 8 * sizeof(unsigned long long), r->stream);
    if (ce == cudaSuccess && mode == HNM_MODE_PATHTRACING)
        ce = cudaFuncSetAttribute(k_isaac_raygen, cudaFuncAttributeMaxDynamicSharedMemorySize, ISAAC_PATHS * 256 * (int)sizeof(uint64_t));
    if (ce == cudaSuccess && mode == HNM_MODE_PATHTRACING)
        ce = cudaFuncSetAttribute(k_isaac_raygen_tm, cudaFuncAttributeMaxDynamicSharedMemorySize, ISAAC_PATHS * 256 * (int)sizeof(uint64_t));
    if (const char* e = getenv("HNM_CARVEOUT")) {
        // experiment: the shared-memory carve-out the kernels that co-reside with k_isaac_raygen ask for
        int c = atoi(e);
        cudaFuncSetAttribute(k_confirm<false>, cudaFuncAttributePreferredSharedMemoryCarveout, c);
        cudaFuncSetAttribute(k_trace<false>, cudaFuncAttributePreferredSharedMemoryCarveout, c);
        cudaFuncSetAttribute(k_trace<true>, cudaFuncAttributePreferredSharedMemoryCarveout, c);
        cudaFuncSetAttribute(k_shade_miss<false>, cudaFuncAttributePreferredSharedMemoryCarveout, c);
        cudaFuncSetAttribute(k_shade_surf<false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, c);
        cudaFuncSetAttribute(k_shade_surf<true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, c);
        cudaFuncSetAttribute(k_nee_resolve<false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, c);
        cudaFuncSetAttribute(k_accumulate, cudaFuncAttributePreferredSharedMemoryCarveout, c);
        cudaFuncSetAttribute(k_batch_begin, cudaFuncAttributePreferredSharedMemoryCarveout, c);
    }
    if (ce != cudaSuccess) { set_error(HNM_ERR_CUDA, std::string("renderer init: ") + cudaGetErrorString(ce)); return bail(HNM_ERR_CUDA); }
    *out = r;
    return 0;
}

int hnm_render_passes(hnm_renderer* r, uint32_t sampling_first, uint32_t count) {
    if (!r) return set_error(HNM_ERR_INVALID, "null renderer");
    if (r->P.mode != HNM_MODE_PATHTRACING && count > 1) count = 1;  // DebugRenderer::max_sampling() == 1
    HNM_CUDA(cudaSetDevice(r->scene->device));
    // The call is cut into ceil(count / max_batch) batches of (nearly) EQUAL size -- 16 passes with room for 3 in flight
    // run as 3 3 3 3 2 2, not 3 3 3 3 3 1: a thin last batch wastes the machine.  The batch after each one is known
    // inside the call; across calls it is guessed as the first batch of an identical call that continues the pass
    // numbering (`sampling` only ever counts up, src/renderer.rs:32).  A wrong guess costs one discarded generation.
    auto batch_size = [&](uint32_t remaining) {
        uint32_t nb = (remaining + r->max_batch - 1) / r->max_batch;
        return (remaining + nb - 1) / nb;
    };
    uint32_t done = 0;
    while (done < count) {
        const uint32_t b = batch_size(count - done);
        const uint32_t left = count - done - b;
        const uint32_t nb = left > 0 ? batch_size(left) : (r->speculate ? batch_size(count) : 0u);
        int rc = run_batch(r, sampling_first + done, b, sampling_first + done + b, nb);
        if (rc) return rc;
        done += b;
    }
    return 0;
}
int hnm_synchronize(hnm_renderer* r) {
    if (!r) return set_error(HNM_ERR_INVALID, "null renderer");
    HNM_CUDA(cudaSetDevice(r->scene->device));
    HNM_CUDA(cudaStreamSynchronize(r->stream));
    return 0;
}
int hnm_clear(hnm_renderer* r) {
    if (!r) return set_error(HNM_ERR_INVALID, "null renderer");
    HNM_CUDA(cudaSetDevice(r->scene->device));
    HNM_CUDA(cudaMemsetAsync(r->P.accum, 0, (size_t)r->padded_rows * r->P.W * 3 * sizeof(double), r->stream));
    HNM_CUDA(cudaMemsetAsync(r->P.stats, 0, S_COUNT * sizeof(unsigned long long), r->stream));
    if (r->P.dbg) HNM_CUDA(cudaMemsetAsync(r->P.dbg, 0, 8 * sizeof(unsigned long long), r->stream));
    r->launches = 0;
    return 0;
}
uint32_t hnm_owned_rows(const hnm_renderer* r) { return r ? r->padded_rows : 0; }
uint32_t hnm_local_row_to_global(const hnm_renderer* r, uint32_t lr) {
    return r ? local_to_global_row_h(lr, r->P.rank, r->P.nranks, r->P.tile_rows) : 0;
}
int hnm_read_accum(hnm_renderer* r, double* rgb) {
    if (!r || !rgb) return set_error(HNM_ERR_INVALID, "null argument");
    HNM_CUDA(cudaSetDevice(r->scene->device));
    HNM_CUDA(cudaMemcpyAsync(rgb, r->P.accum, (size_t)r->padded_rows * r->P.W * 3 * sizeof(double), cudaMemcpyDeviceToHost, r->stream));
    HNM_CUDA(cudaStreamSynchronize(r->stream));
    return 0;
}
int hnm_accum_device_ptr(hnm_renderer* r, void** ptr, size_t* bytes) {
    if (!r || !ptr || !bytes) return set_error(HNM_ERR_INVALID, "null argument");
    *ptr = r->P.accum;
    *bytes = (size_t)r->padded_rows * r->P.W * 3 * sizeof(double);
    return 0;
}

int hnm_deinterleave(hnm_renderer* r, const void* gathered, void* full) {
    if (!r || !gathered || !full) return set_error(HNM_ERR_INVALID, "null argument");
    HNM_CUDA(cudaSetDevice(r->scene->device));
    k_deinterleave<<<r->sm_count * 4, 256, 0, r->stream>>>((const double*)gathered, (double*)full, r->P.W, r->P.H, r->padded_rows, r->P.nranks, r->P.tile_rows);
    HNM_CUDA(cudaGetLastError());
    HNM_CUDA(cudaStreamSynchronize(r->stream));
    return 0;
}

// update_imgbuf on the renderer's stream: tone map + gamma, bilateral, quantise into r->rgb8 (device) and from
// there into the renderer's pinned staging buffer.  Asynchronous; the caller synchronises the stream.
static int enqueue_resolve(hnm_renderer* r, const void* accum_full_device, uint32_t sampling) {
    size_t n = (size_t)r->P.W * r->P.H;
    if (!r->tmp0 || !r->tmp1 || !r->rgb8 || !r->rgb8_host) {
        // all or nothing: a partial failure must not leave a half-initialised set behind for the next call
        double *t0 = nullptr, *t1 = nullptr;
        uint8_t *d8 = nullptr, *h8 = nullptr;
        cudaError_t e = cudaMalloc(&t0, n * 3 * sizeof(double));
        if (e == cudaSuccess) e = cudaMalloc(&t1, n * 3 * sizeof(double));
        if (e == cudaSuccess) e = cudaMalloc(&d8, n * 3);
        if (e == cudaSuccess) e = cudaMallocHost(&h8, n * 3);
        if (e != cudaSuccess) {
            cudaFree(t0); cudaFree(t1); cudaFree(d8);
            if (h8) cudaFreeHost(h8);
            return set_error(HNM_ERR_NOMEM, std::string("resolve buffers: ") + cudaGetErrorString(e));
        }
        r->allocs.push_back(t0); r->allocs.push_back(t1); r->allocs.push_back(d8);
        r->tmp0 = t0; r->tmp1 = t1; r->rgb8 = d8; r->rgb8_host = h8;
    }
    ResolveParams R;
    R.accum = accum_full_device ? (const double*)accum_full_device : r->P.accum;
    R.tmp0 = r->tmp0; R.tmp1 = r->tmp1; R.rgb8 = r->rgb8;
    R.W = r->P.W; R.H = r->P.H;
    // scale = ((sampling * SS * SS) as f64).recip()  (u32 product, src/renderer.rs:65)
    R.scale = 1.0 / (double)(uint32_t)(sampling * r->scene->config.supersampling * r->scene->config.supersampling);
    R.cfg = r->scene->config;
    const int grid = r->sm_count * 8;
    cudaStream_t st = r->stream;
    launch_timed(r, "tonemap_gamma", [&] { k_tonemap_gamma<<<grid, 256, 0, st>>>(R); });
    double *src = r->tmp0, *dst = r->tmp1;
    for (uint32_t it = 0; it < R.cfg.bilateral_iteration; it++) {
        launch_timed(r, "bilateral", [&] { k_bilateral<<<grid, 256, 0, st>>>(R, src, dst); });
        std::swap(src, dst);
    }
    launch_timed(r, "quantise", [&] { k_quantise<<<grid, 256, 0, st>>>(R, src); });
    HNM_CUDA(cudaGetLastError());
    HNM_CUDA(cudaMemcpyAsync(r->rgb8_host, r->rgb8, n * 3, cudaMemcpyDeviceToHost, st));
    return 0;
}

int hnm_resolve(hnm_renderer* r, const void* accum_full_device, uint32_t sampling, uint8_t* rgb8) {
    if (!r || !rgb8) return set_error(HNM_ERR_INVALID, "null argument");
    if (sampling == 0) return set_error(HNM_ERR_STATE, "resolve with sampling == 0");
    if (!accum_full_device && r->P.nranks != 1) return set_error(HNM_ERR_STATE, "a sharded renderer needs the gathered full-image buffer");
    HNM_CUDA(cudaSetDevice(r->scene->device));
    int rc = enqueue_resolve(r, accum_full_device, sampling);
    if (rc) return rc;
    HNM_CUDA(cudaStreamSynchronize(r->stream));
    memcpy(rgb8, r->rgb8_host, (size_t)r->P.W * r->P.H * 3);
    return 0;
}

// The same resolve in two halves, so that a host can enqueue the NEXT passes before it waits for this image
// (`report_progress` every interval, src/renderer.rs:216-226: the image of step i is read while step i + 1 runs).
static int resolve_mark(hnm_renderer* r) {
    if (!r->resolve_done && cudaEventCreateWithFlags(&r->resolve_done, cudaEventDisableTiming) != cudaSuccess)
        return set_error(HNM_ERR_CUDA, "cudaEventCreate failed");
    HNM_CUDA(cudaEventRecord(r->resolve_done, r->stream));
    r->resolve_pending = true;
    return 0;
}
int hnm_resolve_begin(hnm_renderer* r, const void* accum_full_device, uint32_t sampling) {
    if (!r) return set_error(HNM_ERR_INVALID, "null renderer");
    if (sampling == 0) return set_error(HNM_ERR_STATE, "resolve with sampling == 0");
    if (!accum_full_device && r->P.nranks != 1) return set_error(HNM_ERR_STATE, "a sharded renderer needs the gathered full-image buffer");
    if (r->resolve_pending) return set_error(HNM_ERR_STATE, "hnm_resolve_begin: the previous image has not been collected (hnm_resolve_end)");
    HNM_CUDA(cudaSetDevice(r->scene->device));
    int rc = enqueue_resolve(r, accum_full_device, sampling);
    if (rc) return rc;
    return resolve_mark(r);
}
int hnm_resolve_end(hnm_renderer* r, uint8_t* rgb8) {
    if (!r || !rgb8) return set_error(HNM_ERR_INVALID, "null argument");
    if (!r->resolve_pending) return set_error(HNM_ERR_STATE, "hnm_resolve_end without hnm_resolve_begin");
    HNM_CUDA(cudaSetDevice(r->scene->device));
    HNM_CUDA(cudaEventSynchronize(r->resolve_done));
    r->resolve_pending = false;
    memcpy(rgb8, r->rgb8_host, (size_t)r->P.W * r->P.H * 3);
    return 0;
}

int hnm_get_counters(hnm_renderer* r, hnm_counters* out) {
    if (!r || !out) return set_error(HNM_ERR_INVALID, "null argument");
    HNM_CUDA(cudaSetDevice(r->scene->device));
    unsigned long long s[S_COUNT];
    HNM_CUDA(cudaMemcpyAsync(s, r->P.stats, sizeof(s), cudaMemcpyDeviceToHost, r->stream));
    HNM_CUDA(cudaStreamSynchronize(r->stream));
    out->paths = s[S_PATHS]; out->segments = s[S_SEGMENTS]; out->shadow_rays = s[S_SHADOW];
    out->rng_fallbacks = s[S_RNG_FALLBACK]; out->kernel_launches = r->launches;
    out->node_visits = s[S_NODES]; out->prim_tests = s[S_PRIMS]; out->cand_overflows = s[S_OVERFLOW];
    return 0;
}
int hnm_debug_warp_slots(hnm_renderer* r, uint64_t* masks, uint32_t n) {
    if (!r || !masks) return set_error(HNM_ERR_INVALID, "null argument");
    for (uint32_t k = 0; k < n; k++) masks[k] = 0;
    if (!r->P.dbg) return 0;
    HNM_CUDA(cudaSetDevice(r->scene->device));
    unsigned long long m[8];
    HNM_CUDA(cudaMemcpyAsync(m, r->P.dbg, sizeof(m), cudaMemcpyDeviceToHost, r->stream));
    HNM_CUDA(cudaStreamSynchronize(r->stream));
    for (uint32_t k = 0; k < n && k < 8; k++) masks[k] = m[k];
    return 0;
}
// Diagnostics: the per-bounce queue counters of the LAST batch (row b = bounce b, 16 words: rays, miss, delta, nee, events,
// shadow rays, then the work counters)
int hnm_debug_read_counters(hnm_renderer* r, uint32_t* out, uint32_t n) {
    if (!r || !out) return set_error(HNM_ERR_INVALID, "null argument");
    HNM_CUDA(cudaSetDevice(r->scene->device));
    HNM_CUDA(cudaMemcpyAsync(out, r->P.counters, sizeof(uint32_t) * std::min<uint32_t>(n, NUM_COUNTERS), cudaMemcpyDeviceToHost, r->stream));
    HNM_CUDA(cudaStreamSynchronize(r->stream));
    return 0;
}
int hnm_set_precision(hnm_renderer* r, int precision) {
    if (!r) return set_error(HNM_ERR_INVALID, "null renderer");
    if (precision != HNM_PRECISION_EXACT && precision != HNM_PRECISION_FAST_MATH) return set_error(HNM_ERR_INVALID, "bad precision");
    r->fast_math = precision == HNM_PRECISION_FAST_MATH;
    return 0;
}
int hnm_set_profiling(hnm_renderer* r, int enabled) {
    if (!r) return set_error(HNM_ERR_INVALID, "null renderer");
    r->timer.reset();
    r->profiling = enabled != 0;
    return 0;
}
int hnm_mark(hnm_renderer* r, uint32_t slot) {
    if (!r || slot >= 16) return set_error(HNM_ERR_INVALID, "bad mark slot");
    HNM_CUDA(cudaSetDevice(r->scene->device));
    if (!r->marks[slot]) HNM_CUDA(cudaEventCreate(&r->marks[slot]));
    HNM_CUDA(cudaEventRecord(r->marks[slot], r->stream));
    return 0;
}
int hnm_elapsed_ms(hnm_renderer* r, uint32_t a, uint32_t b, float* ms) {
    if (!r || !ms || a >= 16 || b >= 16 || !r->marks[a] || !r->marks[b]) return set_error(HNM_ERR_INVALID, "bad mark slot");
    HNM_CUDA(cudaSetDevice(r->scene->device));
    HNM_CUDA(cudaEventSynchronize(r->marks[a]));
    HNM_CUDA(cudaEventSynchronize(r->marks[b]));
    HNM_CUDA(cudaEventElapsedTime(ms, r->marks[a], r->marks[b]));
    return 0;
}
int hnm_get_kernel_times(hnm_renderer* r, uint32_t max, const char** names, float* ms, uint32_t* launches, uint32_t* n) {
    if (!r || !names || !ms || !launches || !n) return set_error(HNM_ERR_INVALID, "null argument");
    HNM_CUDA(cudaSetDevice(r->scene->device));
    r->timer.collect();
    uint32_t k = 0;
    for (auto& name : r->timer.order) {
        if (k >= max) break;
        auto it = r->timer.totals.find(name);
        names[k] = it->first.c_str();
        ms[k] = (float)it->second.first; launches[k] = it->second.second;
        k++;
    }
    *n = k;
    return 0;
}

// ---- multi-GPU behind the boundary -------------------------------------------------------------------------
// (a) one process, N devices: what a single-process host (the Rust binary) calls instead of hnm_renderer_*.
int hnm_group_create(const hnm_scene_desc* desc, const hnm_camera* camera, uint32_t width, uint32_t height, int mode,
                     uint32_t num_devices, const int* devices, uint32_t tile_rows, uint32_t max_batch, hnm_group** out) {
    if (!desc || !camera || !out || !devices) return set_error(HNM_ERR_INVALID, "null argument");
    *out = nullptr;
    if (num_devices == 0 || num_devices > 64) return set_error(HNM_ERR_INVALID, "bad device count");
    if (tile_rows == 0) tile_rows = 4;
    std::shared_ptr<SceneBuilder> bp;
    int rc = scene_build_cached(desc, bp);  // validated and re-laid out ONCE, uploaded to every device
    if (rc) return rc;
    const SceneBuilder& b = *bp;
    hnm_group* g = new hnm_group();
    g->W = width; g->H = height;
    auto bail = [&](int code) { hnm_group_destroy(g); return code; };
    for (uint32_t k = 0; k < num_devices; k++) {
        hnm_scene* sc = nullptr;
        if ((rc = scene_upload(desc, b, devices[k], &sc))) return bail(rc);
        g->scenes.push_back(sc);
        hnm_shard sh = {k, num_devices, tile_rows, 0};
        hnm_renderer* r = nullptr;
        if ((rc = hnm_renderer_create(sc, camera, width, height, mode, &sh, max_batch, &r))) return bail(rc);
        g->rs.push_back(r);
        cudaEvent_t ev = nullptr;
        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { set_error(HNM_ERR_CUDA, "cudaEventCreate failed"); return bail(HNM_ERR_CUDA); }
        g->copied.push_back(ev);
    }
    // direct NVLink stores into device 0 for the gather (without peer access the copies are staged by the driver)
    for (uint32_t k = 1; k < num_devices; k++) {
        if (devices[k] == devices[0]) continue;
        int can = 0;
        cudaDeviceCanAccessPeer(&can, devices[k], devices[0]);
        if (can) {
            cudaSetDevice(devices[k]);
            cudaError_t e = cudaDeviceEnablePeerAccess(devices[0], 0);
            if (e != cudaSuccess) cudaGetLastError();  // already enabled: fine
        }
    }
    *out = g;
    return 0;
}
void hnm_group_destroy(hnm_group* g) {
    if (!g) return;
    for (auto r : g->rs) hnm_renderer_destroy(r);
    for (auto s : g->scenes) hnm_scene_destroy(s);
    for (auto e : g->copied) if (e) cudaEventDestroy(e);
    delete g;
}
uint32_t hnm_group_size(const hnm_group* g) { return g ? (uint32_t)g->rs.size() : 0; }
hnm_renderer* hnm_group_member(hnm_group* g, uint32_t k) { return (g && k < g->rs.size()) ? g->rs[k] : nullptr; }
int hnm_group_render_passes(hnm_group* g, uint32_t sampling_first, uint32_t count) {
    if (!g) return set_error(HNM_ERR_INVALID, "null group");
    for (auto r : g->rs) { int rc = hnm_render_passes(r, sampling_first, count); if (rc) return rc; }  // asynchronous on every device
    return 0;
}
int hnm_group_synchronize(hnm_group* g) {
    if (!g) return set_error(HNM_ERR_INVALID, "null group");
    for (auto r : g->rs) { int rc = hnm_synchronize(r); if (rc) return rc; }
    return 0;
}
int hnm_group_clear(hnm_group* g) {
    if (!g) return set_error(HNM_ERR_INVALID, "null group");
    for (auto r : g->rs) { int rc = hnm_clear(r); if (rc) return rc; }
    return 0;
}
// shards -> device 0 (peer copies, each on its source device's stream, i.e. after that device's passes), then the
// image in row order on device 0's stream.  Asynchronous.
static int group_gather(hnm_group* g, const double** full) {
    hnm_renderer* r0 = g->rs[0];
    if (g->rs.size() == 1) { *full = r0->P.accum; return 0; }
    HNM_CUDA(cudaSetDevice(r0->scene->device));
    int rc = ensure_gather_buffers(r0);
    if (rc) return rc;
    const size_t shard = (size_t)r0->padded_rows * r0->P.W * 3;
    for (size_t k = 0; k < g->rs.size(); k++) {
        hnm_renderer* r = g->rs[k];
        HNM_CUDA(cudaSetDevice(r->scene->device));
        HNM_CUDA(cudaMemcpyPeerAsync(r0->gathered + k * shard, r0->scene->device, r->P.accum, r->scene->device, shard * sizeof(double), r->stream));
        HNM_CUDA(cudaEventRecord(g->copied[k], r->stream));
    }
    HNM_CUDA(cudaSetDevice(r0->scene->device));
    for (size_t k = 1; k < g->rs.size(); k++) HNM_CUDA(cudaStreamWaitEvent(r0->stream, g->copied[k], 0));
    k_deinterleave<<<r0->sm_count * 4, 256, 0, r0->stream>>>(r0->gathered, r0->full, r0->P.W, r0->P.H, r0->padded_rows, r0->P.nranks, r0->P.tile_rows);
    HNM_CUDA(cudaGetLastError());
    *full = r0->full;
    return 0;
}
int hnm_group_resolve(hnm_group* g, uint32_t sampling, uint8_t* rgb8) {
    if (!g || !rgb8) return set_error(HNM_ERR_INVALID, "null argument");
    if (sampling == 0) return set_error(HNM_ERR_STATE, "resolve with sampling == 0");
    const double* full = nullptr;
    int rc = group_gather(g, &full);
    if (rc) return rc;
    hnm_renderer* r0 = g->rs[0];
    if ((rc = enqueue_resolve(r0, full, sampling))) return rc;   // on device 0 only
    HNM_CUDA(cudaStreamSynchronize(r0->stream));
    memcpy(rgb8, r0->rgb8_host, (size_t)g->W * g->H * 3);
    return 0;
}
int hnm_group_read_accum(hnm_group* g, double* rgb) {
    if (!g || !rgb) return set_error(HNM_ERR_INVALID, "null argument");
    const double* full = nullptr;
    int rc = group_gather(g, &full);
    if (rc) return rc;
    hnm_renderer* r0 = g->rs[0];
    HNM_CUDA(cudaMemcpyAsync(rgb, full, (size_t)g->W * g->H * 3 * sizeof(double), cudaMemcpyDeviceToHost, r0->stream));
    HNM_CUDA(cudaStreamSynchronize(r0->stream));
    return 0;
}
int hnm_group_get_counters(hnm_group* g, hnm_counters* out) {
    if (!g || !out) return set_error(HNM_ERR_INVALID, "null argument");
    memset(out, 0, sizeof(*out));
    for (auto r : g->rs) {
        hnm_counters c;
        int rc = hnm_get_counters(r, &c);
        if (rc) return rc;
        out->paths += c.paths; out->segments += c.segments; out->shadow_rays += c.shadow_rays; out->rng_fallbacks += c.rng_fallbacks;
        out->kernel_launches += c.kernel_launches; out->node_visits += c.node_visits; out->prim_tests += c.prim_tests;
        out->cand_overflows += c.cand_overflows;
    }
    return 0;
}

// (b) one process per device (torchrun / MPI style): the one exchange step of the path, an NCCL all-gather of the f64
// accumulation shards, enqueued on the renderer's own stream right behind its passes.
int hnm_dist_unique_id(uint8_t* id) {
    if (!id) return set_error(HNM_ERR_INVALID, "null argument");
    if (!nccl().ok) return set_error(HNM_ERR_STATE, "NCCL unavailable: " + nccl().why);
    NcclApi::UniqueId u;
    HNM_NCCL(nccl().GetUniqueId(&u));
    memcpy(id, u.internal, HNM_DIST_ID_BYTES);
    return 0;
}
int hnm_dist_init(hnm_renderer* r, const uint8_t* id, uint32_t rank, uint32_t num_ranks) {
    if (!r || !id) return set_error(HNM_ERR_INVALID, "null argument");
    if (rank != r->P.rank || num_ranks != r->P.nranks) return set_error(HNM_ERR_INVALID, "rank / num_ranks differ from the renderer's shard");
    if (!nccl().ok) return set_error(HNM_ERR_STATE, "NCCL unavailable: " + nccl().why);
    if (r->nccl_comm) return set_error(HNM_ERR_STATE, "hnm_dist_init called twice");
    HNM_CUDA(cudaSetDevice(r->scene->device));
    NcclApi::UniqueId u;
    memcpy(u.internal, id, HNM_DIST_ID_BYTES);
    HNM_NCCL(nccl().CommInitRank(&r->nccl_comm, (int)num_ranks, u, (int)rank));
    r->owns_comm = true;
    r->dist_rank = rank; r->dist_nranks = num_ranks;
    return ensure_gather_buffers(r);
}
int hnm_comm_create(int device, const uint8_t* id, uint32_t rank, uint32_t num_ranks, hnm_comm** out) {
    if (!id || !out) return set_error(HNM_ERR_INVALID, "null argument");
    *out = nullptr;
    if (num_ranks == 0 || rank >= num_ranks) return set_error(HNM_ERR_INVALID, "bad rank / num_ranks");
    HNM_CUDA(cudaSetDevice(device));  // (before NCCL is bound: a host without a device never loads it)
    if (!nccl().ok) return set_error(HNM_ERR_STATE, "NCCL unavailable: " + nccl().why);
    NcclApi::UniqueId u;
    memcpy(u.internal, id, HNM_DIST_ID_BYTES);
    void* c = nullptr;
    HNM_NCCL(nccl().CommInitRank(&c, (int)num_ranks, u, (int)rank));
    hnm_comm* h = new hnm_comm();
    h->comm = c; h->device = device; h->rank = rank; h->num_ranks = num_ranks;
    *out = h;
    return 0;
}
void hnm_comm_destroy(hnm_comm* comm) {
    if (!comm) return;
    if (comm->comm && nccl().ok) { cudaSetDevice(comm->device); nccl().CommDestroy(comm->comm); }
    delete comm;
}
int hnm_dist_attach(hnm_renderer* r, hnm_comm* comm) {
    if (!r || !comm) return set_error(HNM_ERR_INVALID, "null argument");
    if (comm->rank != r->P.rank || comm->num_ranks != r->P.nranks) return set_error(HNM_ERR_INVALID, "the communicator's rank / size differ from the renderer's shard");
    if (comm->device != r->scene->device) return set_error(HNM_ERR_INVALID, "the communicator belongs to another device");
    if (r->nccl_comm) return set_error(HNM_ERR_STATE, "the renderer already has a communicator");
    r->nccl_comm = comm->comm;
    r->owns_comm = false;
    r->dist_rank = comm->rank; r->dist_nranks = comm->num_ranks;
    return ensure_gather_buffers(r);
}
static int dist_gather(hnm_renderer* r) {
    if (!r->nccl_comm) return set_error(HNM_ERR_STATE, "hnm_dist_init has not been called");
    HNM_CUDA(cudaSetDevice(r->scene->device));
    const size_t shard = (size_t)r->padded_rows * r->P.W * 3;
    HNM_NCCL(nccl().AllGather(r->P.accum, r->gathered, shard, NCCL_FLOAT64, r->nccl_comm, r->stream));
    k_deinterleave<<<r->sm_count * 4, 256, 0, r->stream>>>(r->gathered, r->full, r->P.W, r->P.H, r->padded_rows, r->P.nranks, r->P.tile_rows);
    HNM_CUDA(cudaGetLastError());
    return 0;
}
// Collective: every rank calls it.  Ranks that pass rgb8 == NULL only take part in the gather (asynchronously);
// a rank that passes a buffer also runs update_imgbuf on the gathered image and returns it.
int hnm_dist_resolve(hnm_renderer* r, uint32_t sampling, uint8_t* rgb8) {
    if (!r) return set_error(HNM_ERR_INVALID, "null renderer");
    if (sampling == 0) return set_error(HNM_ERR_STATE, "resolve with sampling == 0");
    int rc = dist_gather(r);
    if (rc || !rgb8) return rc;
    if ((rc = enqueue_resolve(r, r->full, sampling))) return rc;
    HNM_CUDA(cudaStreamSynchronize(r->stream));
    memcpy(rgb8, r->rgb8_host, (size_t)r->P.W * r->P.H * 3);
    return 0;
}
// Collective, asynchronous: the gather (and, with want_image != 0, update_imgbuf) is enqueued; hnm_resolve_end collects.
int hnm_dist_resolve_begin(hnm_renderer* r, uint32_t sampling, int want_image) {
    if (!r) return set_error(HNM_ERR_INVALID, "null renderer");
    if (sampling == 0) return set_error(HNM_ERR_STATE, "resolve with sampling == 0");
    if (want_image && r->resolve_pending) return set_error(HNM_ERR_STATE, "hnm_dist_resolve_begin: the previous image has not been collected (hnm_resolve_end)");
    int rc = dist_gather(r);
    if (rc || !want_image) return rc;
    if ((rc = enqueue_resolve(r, r->full, sampling))) return rc;
    return resolve_mark(r);
}
int hnm_dist_read_accum(hnm_renderer* r, double* rgb) {
    if (!r) return set_error(HNM_ERR_INVALID, "null renderer");
    int rc = dist_gather(r);
    if (rc || !rgb) return rc;
    HNM_CUDA(cudaMemcpyAsync(rgb, r->full, (size_t)r->P.W * r->P.H * 3 * sizeof(double), cudaMemcpyDeviceToHost, r->stream));
    HNM_CUDA(cudaStreamSynchronize(r->stream));
    return 0;
}

// ---- batch entry points -----------------------------------------------------------------------
int hnm_intersect_batch(hnm_scene* scene, const hnm_ray* rays, uint32_t n, hnm_hit* hits) {
    if (!scene || !rays || !hits) return set_error(HNM_ERR_INVALID, "null argument");
    if (n == 0) return 0;
    HNM_CUDA(cudaSetDevice(scene->device));
    // the production traversal kernel (k_trace) on an SoA copy of the rays, then the material resolve
    std::vector<void*> tmp;
    auto cleanup = [&] { for (auto p : tmp) cudaFree(p); };
    int rc = 0;
    double* ray[6]; double *ht, *hu, *hv; uint2* hid; uint32_t* cnt; hnm_hit* dh; unsigned long long* dstats;
    CandLists cl;
    cl.stride = n;
    if ((rc = dev_alloc(tmp, &cl.id, (size_t)n * TRACE_CAND)) || (rc = dev_alloc(tmp, &cl.lo, (size_t)n * TRACE_CAND)) ||
        (rc = dev_alloc(tmp, &cl.n, (size_t)n)) || (rc = dev_alloc(tmp, &cl.ub, (size_t)n))) { cleanup(); return rc; }
    for (int k = 0; k < 6; k++) if ((rc = dev_alloc(tmp, &ray[k], (size_t)n))) { cleanup(); return rc; }
    if ((rc = dev_alloc(tmp, &ht, (size_t)n)) || (rc = dev_alloc(tmp, &hu, (size_t)n)) || (rc = dev_alloc(tmp, &hv, (size_t)n)) ||
        (rc = dev_alloc(tmp, &hid, (size_t)n)) || (rc = dev_alloc(tmp, &cnt, (size_t)16)) || (rc = dev_alloc(tmp, &dh, (size_t)n)) ||
        (rc = dev_alloc(tmp, &dstats, (size_t)S_COUNT))) { cleanup(); return rc; }
    std::vector<double> soa((size_t)n);
    for (int k = 0; k < 6; k++) {
        for (uint32_t i = 0; i < n; i++) soa[i] = k < 3 ? (&rays[i].origin.x)[k] : (&rays[i].direction.x)[k - 3];
        cudaError_t e = cudaMemcpy(ray[k], soa.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { cleanup(); return set_error(HNM_ERR_CUDA, std::string("cudaMemcpy: ") + cudaGetErrorString(e)); }
    }
    uint32_t hc[16] = {0};
    hc[0] = n;
    cudaMemcpy(cnt, hc, sizeof(hc), cudaMemcpyHostToDevice);
    cudaMemset(dstats, 0, S_COUNT * sizeof(unsigned long long));
    TraceArgs A;
    memset(&A, 0, sizeof(A));
    for (int k = 0; k < 6; k++) A.job[0].ray[k] = ray[k];
    A.job[0].hit_t = ht; A.job[0].hit_u = hu; A.job[0].hit_v = hv; A.job[0].hit_id = hid;
    A.job[0].count = cnt;
    A.njobs = 1;
    A.work = cnt + 1;
    A.work_confirm = cnt + 2;
    A.stats = dstats;
    A.stat_segments = -1; A.stat_nodes = S_NODES; A.stat_prims = S_PRIMS;
    A.cand = cl;
    k_trace<false><<<scene->sm_count * HNM_TRACE_MIN_BLOCKS, TRACE_THREADS>>>(scene->d, A);
    k_confirm<false><<<scene->sm_count * 8, 256>>>(scene->d, A);
    RayPtrs rp;
    for (int k = 0; k < 6; k++) rp.p[k] = ray[k];
    k_hits_to_abi<<<scene->sm_count * 4, 256>>>(scene->d, rp, ht, hu, hv, hid, n, dh);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(hits, dh, (size_t)n * sizeof(hnm_hit), cudaMemcpyDeviceToHost);
    cleanup();
    if (e != cudaSuccess) return set_error(HNM_ERR_CUDA, std::string("intersect_batch: ") + cudaGetErrorString(e));
    return 0;
}

int hnm_isaac64_batch(int device, const uint64_t* seeds, uint32_t n, uint32_t count, uint64_t* out) {
    if (!seeds || !out) return set_error(HNM_ERR_INVALID, "null argument");
    if (n == 0 || count == 0) return 0;
    HNM_CUDA(cudaSetDevice(device));
    int sms = 0, rc = 0;
    if ((rc = device_sm_count(device, &sms))) return rc;
    TmpDev T;
    uint64_t *ds = nullptr, *dout = nullptr;
    if ((rc = T.upload(&ds, seeds, (size_t)n * 4 * sizeof(uint64_t)))) return rc;
    if ((rc = T.alloc(&dout, (size_t)n * count * sizeof(uint64_t)))) return rc;
    if (count <= HNM_RNG_TAIL) {
        size_t smem = (size_t)ISAAC_PATHS * 256 * sizeof(uint64_t);
        const char* e = getenv("HNM_ISAAC_TMEM");
        if (!e || atoi(e) != 0) {
            HNM_CUDA(cudaFuncSetAttribute(k_isaac_batch_tm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_isaac_batch_tm<<<sms, ISAAC_TM_THREADS, smem>>>(ds, n, count, dout);
        } else {
            HNM_CUDA(cudaFuncSetAttribute(k_isaac_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_isaac_batch<<<sms, ISAAC_THREADS, smem>>>(ds, n, count, dout);
        }
    } else {
        k_isaac_full_batch<<<sms, 64>>>(ds, n, count, dout);  // the exact slow path, with refill
    }
    HNM_CUDA(cudaGetLastError());
    HNM_CUDA(cudaMemcpy(out, dout, (size_t)n * count * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return 0;
}
static void default_math_scene(DScene& sc) {
    memset(&sc, 0, sizeof(sc));
    sc.eps = 1e-4; sc.offset = 1e-4; sc.inf = 1e100; sc.gamma = 2.2;
}
int hnm_material_sample_batch(int device, const double* in, uint32_t n, double* out) {
    if (!in || !out) return set_error(HNM_ERR_INVALID, "null argument");
    if (n == 0) return 0;
    HNM_CUDA(cudaSetDevice(device));
    int sms = 0, rc = 0;
    if ((rc = device_sm_count(device, &sms))) return rc;
    TmpDev T;
    double *di = nullptr, *dout = nullptr;
    if ((rc = T.upload(&di, in, (size_t)n * 14 * sizeof(double)))) return rc;
    if ((rc = T.alloc(&dout, (size_t)n * 8 * sizeof(double)))) return rc;
    DScene sc;
    default_math_scene(sc);
    k_material_sample_batch<<<sms * 2, 256>>>(sc, di, n, dout);
    HNM_CUDA(cudaGetLastError());
    HNM_CUDA(cudaMemcpy(out, dout, (size_t)n * 8 * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}
int hnm_material_bsdf_batch(int device, const double* in, uint32_t n, double* out) {
    if (!in || !out) return set_error(HNM_ERR_INVALID, "null argument");
    if (n == 0) return 0;
    HNM_CUDA(cudaSetDevice(device));
    int sms = 0, rc = 0;
    if ((rc = device_sm_count(device, &sms))) return rc;
    TmpDev T;
    double *di = nullptr, *dout = nullptr;
    if ((rc = T.upload(&di, in, (size_t)n * 12 * sizeof(double)))) return rc;
    if ((rc = T.alloc(&dout, (size_t)n * sizeof(double)))) return rc;
    k_material_bsdf_batch<<<sms * 2, 256>>>(di, n, dout);
    HNM_CUDA(cudaGetLastError());
    HNM_CUDA(cudaMemcpy(out, dout, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}
int hnm_math_batch(int device, int fn, const double* x, const double* y, uint32_t n, double* out) {
    if (!x || !y || !out) return set_error(HNM_ERR_INVALID, "null argument");
    if (n == 0) return 0;
    HNM_CUDA(cudaSetDevice(device));
    int sms = 0, rc = 0;
    if ((rc = device_sm_count(device, &sms))) return rc;
    TmpDev T;
    double *dx = nullptr, *dy = nullptr, *dout = nullptr;
    if ((rc = T.upload(&dx, x, (size_t)n * sizeof(double)))) return rc;
    if ((rc = T.upload(&dy, y, (size_t)n * sizeof(double)))) return rc;
    if ((rc = T.alloc(&dout, (size_t)n * sizeof(double)))) return rc;
    k_math_batch<<<sms * 2, 256>>>(fn, dx, dy, n, dout);
    HNM_CUDA(cudaGetLastError());
    HNM_CUDA(cudaMemcpy(out, dout, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}
// `Texture::sample` (src/texture.rs:108-114: bilinear on the gamma-space texels, v flip, u32 wrap at the top row,
// pow 2.2, times the tint) on n independent (u, v) pairs; image < 0 = a constant-colour texture
int hnm_texture_sample_batch(hnm_scene* scene, int32_t image, const double* tint, const double* uv, uint32_t n, double* rgb) {
    if (!scene || !tint || !uv || !rgb) return set_error(HNM_ERR_INVALID, "null argument");
    if (image >= (int32_t)scene->num_images) return set_error(HNM_ERR_INVALID, "bad image index");
    if (n == 0) return 0;
    HNM_CUDA(cudaSetDevice(scene->device));
    TmpDev T;
    int rc = 0;
    double *duv = nullptr, *dout = nullptr;
    if ((rc = T.upload(&duv, uv, (size_t)n * 2 * sizeof(double)))) return rc;
    if ((rc = T.alloc(&dout, (size_t)n * 3 * sizeof(double)))) return rc;
    DTexture t;
    t.r = tint[0]; t.g = tint[1]; t.b = tint[2]; t.image = image < 0 ? -1 : image; t._pad = 0;
    k_texture_sample_batch<<<scene->sm_count * 2, 256>>>(scene->d, t, duv, n, dout);
    HNM_CUDA(cudaGetLastError());
    HNM_CUDA(cudaMemcpy(rgb, dout, (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}
// `Skybox::sample` (src/scene.rs:295-319) on n directions (xyz triples, not normalised by the callee)
int hnm_skybox_sample_batch(hnm_scene* scene, const double* directions, uint32_t n, double* rgb) {
    if (!scene || !directions || !rgb) return set_error(HNM_ERR_INVALID, "null argument");
    if (n == 0) return 0;
    HNM_CUDA(cudaSetDevice(scene->device));
    TmpDev T;
    int rc = 0;
    double *dd = nullptr, *dout = nullptr;
    if ((rc = T.upload(&dd, directions, (size_t)n * 3 * sizeof(double)))) return rc;
    if ((rc = T.alloc(&dout, (size_t)n * 3 * sizeof(double)))) return rc;
    k_skybox_sample_batch<<<scene->sm_count * 2, 256>>>(scene->d, dd, n, dout);
    HNM_CUDA(cudaGetLastError());
    HNM_CUDA(cudaMemcpy(rgb, dout, (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

}  // extern "C"
