// hanamaru_b200.cu -- wavefront radiance-loop kernels for sm_100a and the C ABI of
// include/hanamaru_b200.h.  Compile with -fmad=false (parity: no FMA contraction).
//
// Kernel set for one batch of passes (PathTracingRenderer, src/renderer.rs:148-203):
//   k_isaac_raygen   ISAAC-64 seeding per path (rand 0.4 StdRng; 2 KB state per path in shared
//                    memory, 112 paths per CTA) + thin-lens camera ray (src/camera.rs:66-96)
//   k_rng_overflow   exact slow path for the (rare) paths whose lens rejection loop outruns the
//                    stored tail of the random stream
//   per bounce b = 1 .. bounce_limit-1:
//     k_extend       closest hit through the unified BVH (hnm_device.cuh: trace); classifies the
//                    hit into miss / delta-BSDF / NEE-BSDF queues (material sort, warp-aggregated)
//     k_shade_miss   Skybox::sample (src/scene.rs:295-319), radiance update, path ends
//     k_shade_surf   material resolve, BSDF sample, NEE shadow ray for Diffuse/GGX
//                    (src/renderer.rs:269-296), throughput update, compaction into the next queue
//   k_accumulate     per pixel: sum of the 4 sub-pixel paths in the reference's order, += into
//                    the f64 accumulation buffer (src/renderer.rs:37,56)
// Resolve (src/renderer.rs:64-90): k_tonemap_gamma -> k_bilateral -> k_quantise.
// Queue sizes live in device memory; a batch is enqueued without any host synchronisation.
#include <algorithm>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

#include "hnm_device.cuh"
#include "hnm_scene.cuh"

namespace hnm {
thread_local std::string g_last_error;

constexpr int RNG_TAIL = HNM_RNG_TAIL;  // u64 outputs kept per path
constexpr int ISAAC_THREADS = 112;      // 112 x 2 KB = 224 KB of the 227 KB a CTA may use
constexpr int MAX_BOUNCE = 64;

// counters[]: per bounce b (1-origin) four queue lengths
enum { C_RAY = 0, C_MISS = 1, C_DELTA = 2, C_NEE = 3, C_STRIDE = 4 };
// stats[] (u64)
enum { S_PATHS = 0, S_SEGMENTS = 1, S_SHADOW = 2, S_RNG_FALLBACK = 3, S_NODES = 4, S_PRIMS = 5, S_COUNT = 8 };

struct RParams {
    DScene sc;
    hnm_camera cam;
    uint32_t W, H, ss, spp;
    uint32_t npix;        // owned pixels that exist in the image
    uint32_t real_rows;   // owned rows that exist
    uint32_t rank, nranks, tile_rows;
    uint32_t batch, sampling_first;
    uint32_t N;           // paths in this batch = batch * npix * spp
    uint32_t cap;         // allocated paths
    int mode;
    int tail_k;           // usable words of the RNG tail (<= RNG_TAIL; smaller only in tests)
    double* ray[2][6];
    double* thr[2][3];
    uint32_t* pid[2];
    double* L[3];
    uint8_t* cursor;
    uint64_t* tail;       // [RNG_TAIL][cap]
    double* hit_t; double* hit_u; double* hit_v; uint2* hit_id;
    uint32_t* q_miss; uint32_t* q_delta; uint32_t* q_nee; uint32_t* q_ovf;
    uint32_t* counters;   // [(MAX_BOUNCE+2) * C_STRIDE] + overflow count at the end
    unsigned long long* stats;
    double* accum;        // [padded_rows * W * 3]
};
constexpr int OVF_COUNTER = (MAX_BOUNCE + 2) * C_STRIDE;

// image row of local row lr (interleaved row tiles)
HNM_D uint32_t local_to_global_row(const RParams& P, uint32_t lr) {
    uint32_t lt = lr / P.tile_rows;
    return (lt * P.nranks + P.rank) * P.tile_rows + (lr % P.tile_rows);
}
__host__ __device__ inline uint32_t local_to_global_row_h(uint32_t lr, uint32_t rank, uint32_t nranks, uint32_t tile_rows) {
    uint32_t lt = lr / tile_rows;
    return (lt * nranks + rank) * tile_rows + (lr % tile_rows);
}

// path p -> pass, local pixel, sub-pixel; and the normalized coordinate of src/renderer.rs:34-36,51-54
struct PathCoord {
    uint32_t pass, pix, sub, x, y;
    double ncx, ncy;
};
HNM_D PathCoord path_coord(const RParams& P, uint32_t p) {
    PathCoord c;
    c.sub = p % P.spp;
    uint32_t r = p / P.spp;
    c.pix = r % P.npix;
    c.pass = r / P.npix;
    uint32_t lr = c.pix / P.W;
    c.x = c.pix - lr * P.W;
    c.y = local_to_global_row(P, lr);
    uint32_t sx = c.sub % P.ss, sy = c.sub / P.ss;
    double fx = (double)c.x, fy = (double)(P.H - c.y);  // frag_coord = (x, height - y)
    double offx = (double)sx / (double)P.ss - 0.5, offy = (double)sy / (double)P.ss - 0.5;
    double rx = (double)P.W, ry = (double)P.H;
    double m = fmin(rx, ry);
    c.ncx = ((fx + offx) * 2.0 - rx) / m;
    c.ncy = ((fy + offy) * 2.0 - ry) / m;
    return c;
}

// ------------------------------------------------------------------------------------ ISAAC-64
// rand 0.4.3 src/prng/isaac64.rs (third-party, restated; pinned by rand's own KATs in the tests).
#define ISAAC_MIX(a, b, c, d, e, f, g, h) \
    a -= e; f ^= h >> 9;  h += a;         \
    b -= f; g ^= a << 9;  a += b;         \
    c -= g; h ^= b >> 23; b += c;         \
    d -= h; a ^= c << 15; c += d;         \
    e -= a; b ^= d >> 14; d += e;         \
    f -= b; c ^= e << 20; e += f;         \
    g -= c; d ^= f >> 17; f += g;         \
    h -= d; e ^= g << 14; g += h;

// `mem` is this thread's column of a [256][T] u64 array (shared memory: conflict-free for any
// per-lane index because the bank depends only on the lane).  Outputs rsl[i] are handed to `sink`.
template <int T, typename Sink>
__device__ __forceinline__ void isaac64_seed(uint64_t* mem, uint64_t s0, uint64_t s1, uint64_t s2, uint64_t s3, Sink sink) {
#define MEM(i) mem[(i) * T]
    uint64_t a, b, c, d, e, f, g, h;
    a = b = c = d = e = f = g = h = 0x9e3779b97f4a7c13ull;
#pragma unroll 1
    for (int i = 0; i < 4; i++) { ISAAC_MIX(a, b, c, d, e, f, g, h) }
    // first pass mixes in rsl = [s0 s1 s2 s3 0 0 ...]
    a += s0; b += s1; c += s2; d += s3;
#pragma unroll 1
    for (int i = 0; i < 256; i += 8) {
        ISAAC_MIX(a, b, c, d, e, f, g, h)
        MEM(i) = a; MEM(i + 1) = b; MEM(i + 2) = c; MEM(i + 3) = d;
        MEM(i + 4) = e; MEM(i + 5) = f; MEM(i + 6) = g; MEM(i + 7) = h;
    }
    // second pass mixes in mem
#pragma unroll 1
    for (int i = 0; i < 256; i += 8) {
        a += MEM(i); b += MEM(i + 1); c += MEM(i + 2); d += MEM(i + 3);
        e += MEM(i + 4); f += MEM(i + 5); g += MEM(i + 6); h += MEM(i + 7);
        ISAAC_MIX(a, b, c, d, e, f, g, h)
        MEM(i) = a; MEM(i + 1) = b; MEM(i + 2) = c; MEM(i + 3) = d;
        MEM(i + 4) = e; MEM(i + 5) = f; MEM(i + 6) = g; MEM(i + 7) = h;
    }
    // isaac64(): a = 0, b = 0, c = 1  ->  aa = 0, bb = 1
    uint64_t aa = 0, bb = 1;
#define ISAAC_STEP(mixexpr, i, i2)                              \
    {                                                           \
        uint64_t x = MEM(i);                                    \
        aa = (mixexpr) + MEM(i2);                               \
        uint64_t y = MEM(((uint32_t)x >> 3) & 255u) + aa + bb;  \
        MEM(i) = y;                                             \
        bb = MEM(((uint32_t)y >> 11) & 255u) + x;               \
        sink(i, bb);                                            \
    }
#pragma unroll 1
    for (int base = 0; base < 128; base += 4) {
        ISAAC_STEP(~(aa ^ (aa << 21)), base, base + 128)
        ISAAC_STEP(aa ^ (aa >> 5), base + 1, base + 129)
        ISAAC_STEP(aa ^ (aa << 12), base + 2, base + 130)
        ISAAC_STEP(aa ^ (aa >> 33), base + 3, base + 131)
    }
#pragma unroll 1
    for (int base = 128; base < 256; base += 4) {
        ISAAC_STEP(~(aa ^ (aa << 21)), base, base - 128)
        ISAAC_STEP(aa ^ (aa >> 5), base + 1, base - 127)
        ISAAC_STEP(aa ^ (aa << 12), base + 2, base - 126)
        ISAAC_STEP(aa ^ (aa >> 33), base + 3, base - 125)
    }
#undef ISAAC_STEP
#undef MEM
}

// Complete generator with refill, state in local memory: the exact slow path.
struct IsaacFull {
    uint64_t rsl[256], mem[256];
    uint64_t a, b, c;
    uint32_t cnt;
    __device__ void round() {
        c += 1;
        uint64_t aa = a, bb = b + c;
        for (int half = 0; half < 2; half++) {
            int mr = half == 0 ? 0 : 128, m2 = half == 0 ? 128 : 0;
            for (int base = 0; base < 128; base += 4) {
                for (int j = 0; j < 4; j++) {
                    uint64_t mixv = j == 0 ? ~(aa ^ (aa << 21)) : j == 1 ? (aa ^ (aa >> 5)) : j == 2 ? (aa ^ (aa << 12)) : (aa ^ (aa >> 33));
                    uint64_t x = mem[base + j + mr];
                    aa = mixv + mem[base + j + m2];
                    uint64_t y = mem[(x >> 3) & 255] + aa + bb;
                    mem[base + j + mr] = y;
                    bb = mem[(y >> 11) & 255] + x;
                    rsl[base + j + mr] = bb;
                }
            }
        }
        a = aa; b = bb; cnt = 256;
    }
    __device__ void seed(uint64_t s0, uint64_t s1, uint64_t s2, uint64_t s3) {
        for (int i = 0; i < 256; i++) rsl[i] = 0;
        rsl[0] = s0; rsl[1] = s1; rsl[2] = s2; rsl[3] = s3;
        a = b = c = 0;
        uint64_t a_, b_, c_, d_, e_, f_, g_, h_;
        a_ = b_ = c_ = d_ = e_ = f_ = g_ = h_ = 0x9e3779b97f4a7c13ull;
        for (int i = 0; i < 4; i++) { ISAAC_MIX(a_, b_, c_, d_, e_, f_, g_, h_) }
        for (int pass = 0; pass < 2; pass++) {
            const uint64_t* src = pass == 0 ? rsl : mem;
            for (int i = 0; i < 256; i += 8) {
                a_ += src[i]; b_ += src[i + 1]; c_ += src[i + 2]; d_ += src[i + 3];
                e_ += src[i + 4]; f_ += src[i + 5]; g_ += src[i + 6]; h_ += src[i + 7];
                ISAAC_MIX(a_, b_, c_, d_, e_, f_, g_, h_)
                mem[i] = a_; mem[i + 1] = b_; mem[i + 2] = c_; mem[i + 3] = d_;
                mem[i + 4] = e_; mem[i + 5] = f_; mem[i + 6] = g_; mem[i + 7] = h_;
            }
        }
        round();
    }
    __device__ uint64_t next_u64() {
        if (cnt == 0) round();
        cnt -= 1;
        return rsl[cnt & 255];
    }
};

HNM_D void path_seed(const PathCoord& c, uint32_t sampling, uint64_t& s0, uint64_t& s1, uint64_t& s2, uint64_t& s3) {
    // src/renderer.rs:165-167
    s0 = 8700304ull;
    s1 = (uint64_t)sampling;
    s2 = f64_as_u64((4.0 + c.ncx) * 100870.0);
    s3 = f64_as_u64((4.0 + c.ncy) * 100304.0);
}

// thin-lens ray from an accepted lens sample (src/camera.rs:83-96)
HNM_D void lens_ray(const hnm_camera& cm, double ncx, double ncy, double sqx, double sqy, D3& origin, D3& direction) {
    double lx = sqx * cm.lens_radius, ly = sqy * cm.lens_radius;
    D3 lens_pos = d3(cm.right) * lx + d3(cm.up) * ly;
    origin = d3(cm.eye) + lens_pos;
    direction = normalize(ncx * d3(cm.plane_half_right) + ncy * d3(cm.plane_half_up) + cm.focus_distance * d3(cm.forward) - lens_pos);
}

HNM_D void store_ray(const RParams& P, int buf, uint32_t q, D3 o, D3 d, D3 t, uint32_t pid) {
    P.ray[buf][0][q] = o.x; P.ray[buf][1][q] = o.y; P.ray[buf][2][q] = o.z;
    P.ray[buf][3][q] = d.x; P.ray[buf][4][q] = d.y; P.ray[buf][5][q] = d.z;
    P.thr[buf][0][q] = t.x; P.thr[buf][1][q] = t.y; P.thr[buf][2][q] = t.z;
    P.pid[buf][q] = pid;
}

__global__ void __launch_bounds__(ISAAC_THREADS, 1) k_isaac_raygen(RParams P) {
    extern __shared__ uint64_t smem_isaac[];
    const int tid = threadIdx.x;
    uint64_t* mem = smem_isaac + tid;
    const uint32_t N = P.N, cap = P.cap;
    for (uint32_t p = blockIdx.x * ISAAC_THREADS + tid; p < N; p += gridDim.x * ISAAC_THREADS) {
        PathCoord c = path_coord(P, p);
        uint64_t s0, s1, s2, s3;
        path_seed(c, P.sampling_first + c.pass, s0, s1, s2, s3);
        uint64_t* tail = P.tail + p;
        // outputs are consumed from rsl[255] downwards: word j of the stream = rsl[255 - j]
        isaac64_seed<ISAAC_THREADS>(mem, s0, s1, s2, s3, [&](int i, uint64_t v) {
            if (i >= 256 - RNG_TAIL) tail[(size_t)(255 - i) * cap] = v;
        });
        // sample_on_lens (src/camera.rs:66-81): rejection loop over pairs of the stream
        int cur = 0;
        double sqx = 0.0, sqy = 0.0;
        bool ok = false;
        while (cur + 2 <= P.tail_k) {
            double u = u64_to_f64(tail[(size_t)cur * cap]);
            double v = u64_to_f64(tail[(size_t)(cur + 1) * cap]);
            cur += 2;
            sqx = 2.0 * u - 1.0;
            sqy = 2.0 * v - 1.0;
            if (P.cam.lens_shape == 0 || sqx * sqx + sqy * sqy < 1.0) { ok = true; break; }
        }
        P.L[0][p] = 0.0; P.L[1][p] = 0.0; P.L[2][p] = 0.0;
        if (!ok || cur + 2 * (int)(P.sc.bounce_limit - 1) > P.tail_k) {
            // the stored tail is too short for this path: exact slow path (k_rng_overflow)
            uint32_t slot = atomicAdd(&P.counters[OVF_COUNTER], 1u);
            P.q_ovf[slot] = p;
            P.pid[0][p] = 0xFFFFFFFFu;  // parked until the overflow kernel fills it in
            continue;
        }
        D3 o, d;
        lens_ray(P.cam, c.ncx, c.ncy, sqx, sqy, o, d);
        store_ray(P, 0, p, o, d, splat(1.0), p);
        P.cursor[p] = (uint8_t)cur;
    }
    if (blockIdx.x == 0 && tid == 0) {
        P.counters[1 * C_STRIDE + C_RAY] = N;
        atomicAdd(&P.stats[S_PATHS], (unsigned long long)N);
    }
}

__global__ void k_rng_overflow(RParams P) {
    uint32_t n = P.counters[OVF_COUNTER];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t p = P.q_ovf[i];
        PathCoord c = path_coord(P, p);
        uint64_t s0, s1, s2, s3;
        path_seed(c, P.sampling_first + c.pass, s0, s1, s2, s3);
        IsaacFull rng;
        rng.seed(s0, s1, s2, s3);
        double sqx, sqy;
        for (;;) {
            double u = u64_to_f64(rng.next_u64());
            double v = u64_to_f64(rng.next_u64());
            sqx = 2.0 * u - 1.0;
            sqy = 2.0 * v - 1.0;
            if (P.cam.lens_shape == 0 || sqx * sqx + sqy * sqy < 1.0) break;
        }
        // the per-bounce pairs follow; park them at the start of this path's tail
        int need = 2 * (int)(P.sc.bounce_limit - 1);
        for (int j = 0; j < need && j < RNG_TAIL; j++) P.tail[(size_t)j * P.cap + p] = rng.next_u64();
        D3 o, d;
        lens_ray(P.cam, c.ncx, c.ncy, sqx, sqy, o, d);
        store_ray(P, 0, p, o, d, splat(1.0), p);
        P.cursor[p] = 0;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&P.stats[S_RNG_FALLBACK], (unsigned long long)n);
}

// DebugRenderer: pinhole ray, no RNG (src/camera.rs:98-107, src/renderer.rs:117)
__global__ void k_raygen_debug(RParams P) {
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < P.N; p += gridDim.x * blockDim.x) {
        PathCoord c = path_coord(P, p);
        D3 o = d3(P.cam.eye);
        D3 d = normalize(c.ncx * d3(P.cam.plane_half_right) + c.ncy * d3(P.cam.plane_half_up) + P.cam.focus_distance * d3(P.cam.forward));
        store_ray(P, 0, p, o, d, splat(1.0), p);
        P.L[0][p] = 0.0; P.L[1][p] = 0.0; P.L[2][p] = 0.0;
        P.cursor[p] = 0;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        P.counters[1 * C_STRIDE + C_RAY] = P.N;
        atomicAdd(&P.stats[S_PATHS], (unsigned long long)P.N);
    }
}

// append `value` to a queue if `pred`; one atomic per warp
HNM_D void queue_push(bool pred, uint32_t* counter, uint32_t* queue, uint32_t value) {
    unsigned mask = __ballot_sync(__activemask(), pred);
    if (!pred) return;
    int lane = threadIdx.x & 31;
    int leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(mask));
    base = __shfl_sync(mask, base, leader);
    queue[base + __popc(mask & ((1u << lane) - 1u))] = value;
}
HNM_D uint32_t queue_alloc(bool pred, uint32_t* counter) {
    unsigned mask = __ballot_sync(__activemask(), pred);
    if (!pred) return 0;
    int lane = threadIdx.x & 31;
    int leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(mask));
    base = __shfl_sync(mask, base, leader);
    return base + __popc(mask & ((1u << lane) - 1u));
}

HNM_D D3 load_ray_o(const RParams& P, int buf, uint32_t q) { return d3(P.ray[buf][0][q], P.ray[buf][1][q], P.ray[buf][2][q]); }
HNM_D D3 load_ray_d(const RParams& P, int buf, uint32_t q) { return d3(P.ray[buf][3][q], P.ray[buf][4][q], P.ray[buf][5][q]); }
HNM_D D3 load_thr(const RParams& P, int buf, uint32_t q) { return d3(P.thr[buf][0][q], P.thr[buf][1][q], P.thr[buf][2][q]); }

// ------------------------------------------------------------------------------------ extend
template <bool STATS>
__global__ void __launch_bounds__(256) k_extend(RParams P, int bounce, int buf, int classify) {
    const uint32_t n = P.counters[bounce * C_STRIDE + C_RAY];
    TraceStats st;
    st.nodes = 0; st.prims = 0;
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t n_round = (n + 31u) & ~31u;  // keep warps converged for the ballots
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n_round; q += stride) {
        bool active = q < n && P.pid[buf][q] != 0xFFFFFFFFu;
        int cls = -1;
        if (active) {
            D3 o = load_ray_o(P, buf, q), d = load_ray_d(P, buf, q);
            Hit h = trace<STATS>(P.sc, o, d, &st);
            P.hit_t[q] = h.t; P.hit_u[q] = h.u; P.hit_v[q] = h.v;
            P.hit_id[q] = make_uint2(h.kind, h.id);
            if (h.kind == LEAF_NONE) cls = C_MISS;
            else {
                uint32_t el = h.kind == LEAF_TRI ? P.sc.tri_elem[h.id] : h.id;
                int surface = P.sc.materials[P.sc.elements[el].material].surface;
                cls = nee_available(surface) ? C_NEE : C_DELTA;
            }
        }
        if (classify) {
            queue_push(cls == C_MISS, &P.counters[bounce * C_STRIDE + C_MISS], P.q_miss, q);
            queue_push(cls == C_DELTA, &P.counters[bounce * C_STRIDE + C_DELTA], P.q_delta, q);
            queue_push(cls == C_NEE, &P.counters[bounce * C_STRIDE + C_NEE], P.q_nee, q);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&P.stats[S_SEGMENTS], (unsigned long long)n);
    if (STATS) {
        atomicAdd(&P.stats[S_NODES], (unsigned long long)st.nodes);
        atomicAdd(&P.stats[S_PRIMS], (unsigned long long)st.prims);
    }
}

// ------------------------------------------------------------------------------------ shade
HNM_D Rand2 bounce_random(const RParams& P, uint32_t pid, int bounce) {
    // `let random = rng.gen::<(f64, f64)>()` at the top of every bounce (src/renderer.rs:175)
    size_t w = (size_t)P.cursor[pid] + 2u * (uint32_t)(bounce - 1);
    Rand2 r;
    r.r0 = u64_to_f64(P.tail[w * P.cap + pid]);
    r.r1 = u64_to_f64(P.tail[(w + 1) * P.cap + pid]);
    return r;
}

__global__ void __launch_bounds__(256) k_shade_miss(RParams P, int bounce, int buf) {
    const uint32_t n = P.counters[bounce * C_STRIDE + C_MISS];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t q = P.q_miss[i];
        uint32_t pid = P.pid[buf][q];
        D3 d = load_ray_d(P, buf, q);
        D3 thr = load_thr(P, buf, q);
        D3 emission = skybox_sample(P.sc, d);  // src/scene.rs:398
        // accumulation += reflectance * emission (src/renderer.rs:196); the path ends (!hit, :199)
        D3 L = d3(P.L[0][pid], P.L[1][pid], P.L[2][pid]);
        L = L + thr * emission;
        P.L[0][pid] = L.x; P.L[1][pid] = L.y; P.L[2][pid] = L.z;
    }
}

// src/renderer.rs:269-296 for one shading point; traces one closest-hit shadow ray per emitter
template <bool STATS>
HNM_D D3 next_event_estimation(const RParams& P, Rand2 random, double cos_phi, double sin_phi, D3 position, D3 view, D3 normal,
                               const PointMaterial& material, TraceStats* st) {
    const DScene& sc = P.sc;
    D3 accumulation = splat(0.0);
    for (uint32_t k = 0; k < sc.num_emissions; k++) {
        const DElement& e = sc.elements[sc.emissions[k]];
        // Sphere::sample_on_surface (src/scene.rs:92-101); theta = PI2 * random.0 = phi of the BSDF sample
        double unit_z = 1.0 - 2.0 * random.r1;
        double a = __dsqrt_rn(1.0 - unit_z * unit_z);
        D3 s_normal = d3(a * cos_phi, a * sin_phi, unit_z);
        D3 s_position = d3(e.ax, e.ay, e.az) + (e.radius + sc.offset) * s_normal;
        double pdf = 1.0 / (4.0 * HNM_PI * e.radius * e.radius);
        D3 shadow_vec = s_position - position;
        D3 shadow_dir = normalize(shadow_vec);
        Hit h = trace<STATS>(sc, position, shadow_dir, st);
        if (h.kind != LEAF_NONE) {
            D3 hit_pos = position + shadow_dir * h.t;
            if (norm(hit_pos - s_position) < sc.offset * 4.0) {  // Vector3::approximately (src/vector.rs:89-91)
                uint32_t el = h.kind == LEAF_TRI ? sc.tri_elem[h.id] : h.id;
                const DMaterial& hm = sc.materials[sc.elements[el].material];
                D3 emission;
                if (hm.emission.image >= 0) {
                    SurfacePoint sp = surface_point(sc, h, position, shadow_dir, true);
                    emission = texture_sample(sc, hm.emission, sp.u, sp.v);
                } else {
                    emission = d3(hm.emission.r, hm.emission.g, hm.emission.b);
                }
                double dot_0 = fabs(dot(normal, shadow_dir));
                double dot_l = fabs(dot(s_normal, shadow_dir));
                double distance_pow2 = dot(shadow_vec, shadow_vec);
                double g = (dot_0 * dot_l) / distance_pow2;
                accumulation = accumulation + emission * bsdf(material, view, normal, shadow_dir) * g / pdf;
            }
        }
    }
    return accumulation * material.albedo;
}

// One surface interaction of PathTracingRenderer::calc_pixel (src/renderer.rs:176-199) for queue
// `cls` (C_DELTA: Specular / Refraction / GGXRefraction, C_NEE: Diffuse / GGX).
template <bool NEE, bool STATS>
__global__ void __launch_bounds__(256) k_shade_surf(RParams P, int bounce, int buf) {
    const int cls = NEE ? C_NEE : C_DELTA;
    const uint32_t n = P.counters[bounce * C_STRIDE + cls];
    const uint32_t* queue = NEE ? P.q_nee : P.q_delta;
    const bool last_bounce = (uint32_t)bounce + 1 >= P.sc.bounce_limit;
    TraceStats st;
    st.nodes = 0; st.prims = 0;
    uint32_t shadow_rays = 0;
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t n_round = (n + 31u) & ~31u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        bool alive = false;
        D3 no, nd, nthr;
        uint32_t pid = 0;
        if (i < n) {
            uint32_t q = queue[i];
            pid = P.pid[buf][q];
            D3 o = load_ray_o(P, buf, q), d = load_ray_d(P, buf, q);
            D3 thr = load_thr(P, buf, q);
            Hit h;
            h.t = P.hit_t[q]; h.u = P.hit_u[q]; h.v = P.hit_v[q];
            uint2 hid = P.hit_id[q];
            h.kind = hid.x; h.id = hid.y;
            uint32_t el = h.kind == LEAF_TRI ? P.sc.tri_elem[h.id] : h.id;
            const DMaterial& dm_ = P.sc.materials[P.sc.elements[el].material];
            SurfacePoint sp = surface_point(P.sc, h, o, d, dm_.has_image != 0);
            PointMaterial pm = resolve_material(P.sc, dm_, sp.u, sp.v);
            Rand2 random = bounce_random(P, pid, bounce);
            double cos_phi = 1.0, sin_phi = 0.0;
            if (pm.surface != HNM_SURFACE_SPECULAR && pm.surface != HNM_SURFACE_REFRACTION) dm::sincos(HNM_PI2 * random.r0, sin_phi, cos_phi);
            D3 view = -d;
            SampleResult res;
            bool some = material_sample(P.sc, pm, random, cos_phi, sin_phi, sp.position, view, sp.normal, res);
            if (some) {
                D3 L = d3(P.L[0][pid], P.L[1][pid], P.L[2][pid]);
                if (NEE) {
                    shadow_rays += P.sc.num_emissions;
                    D3 nee = next_event_estimation<STATS>(P, random, cos_phi, sin_phi, res.origin, view, sp.normal, pm, &st);
                    L = L + thr * nee;  // src/renderer.rs:183
                }
                L = L + thr * pm.emission;                          // :196
                nthr = thr * (pm.albedo * res.reflectance);          // :197
                P.L[0][pid] = L.x; P.L[1][pid] = L.y; P.L[2][pid] = L.z;
                alive = !all_zero(nthr) && !last_bounce;             // :199 and the loop bound :174
                no = res.origin; nd = res.direction;
            }
            // None: `break` before the emission is added (src/renderer.rs:190-193)
        }
        uint32_t q2 = queue_alloc(alive, &P.counters[(bounce + 1) * C_STRIDE + C_RAY]);
        if (alive) store_ray(P, buf ^ 1, q2, no, nd, nthr, pid);
    }
    if (NEE) {
        for (int o = 16; o > 0; o >>= 1) shadow_rays += __shfl_xor_sync(0xFFFFFFFFu, shadow_rays, o);
        if ((threadIdx.x & 31) == 0 && shadow_rays) atomicAdd(&P.stats[S_SHADOW], (unsigned long long)shadow_rays);
    }
    if (STATS && NEE) {
        atomicAdd(&P.stats[S_NODES], (unsigned long long)st.nodes);
        atomicAdd(&P.stats[S_PRIMS], (unsigned long long)st.prims);
    }
}

// DebugRenderer::calc_pixel (src/renderer.rs:116-139)
__global__ void __launch_bounds__(256) k_debug_shade(RParams P) {
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < P.N; q += gridDim.x * blockDim.x) {
        D3 o = load_ray_o(P, 0, q), d = load_ray_d(P, 0, q);
        Hit h;
        h.t = P.hit_t[q]; h.u = P.hit_u[q]; h.v = P.hit_v[q];
        uint2 hid = P.hit_id[q];
        h.kind = hid.x; h.id = hid.y;
        D3 color;
        if (h.kind == LEAF_NONE) {
            color = skybox_sample(P.sc, d);
        } else {
            uint32_t el = h.kind == LEAF_TRI ? P.sc.tri_elem[h.id] : h.id;
            const DMaterial& dm_ = P.sc.materials[P.sc.elements[el].material];
            SurfacePoint sp = surface_point(P.sc, h, o, d, dm_.has_image != 0);
            if (P.mode == HNM_MODE_DEBUG_SHADING) {
                PointMaterial pm = resolve_material(P.sc, dm_, sp.u, sp.v);
                D3 light_direction = normalize(d3(1.0, 2.0, -1.0));
                Hit sh = trace<false>(P.sc, sp.position + sp.normal * P.sc.offset, light_direction, nullptr);
                double shadow = sh.kind != LEAF_NONE ? 0.5 : 1.0;
                double diffuse = fmax(dot(sp.normal, light_direction), 0.0);
                color = pm.emission + pm.albedo * diffuse * shadow;
            } else if (P.mode == HNM_MODE_DEBUG_NORMAL) {
                color = sp.normal;
            } else if (P.mode == HNM_MODE_DEBUG_DEPTH) {
                color = splat(0.5 * h.t / P.cam.focus_distance);
            } else {
                color = splat(fabs(h.t - P.cam.focus_distance));
            }
        }
        P.L[0][q] = color.x; P.L[1][q] = color.y; P.L[2][q] = color.z;
    }
    if (P.mode == HNM_MODE_DEBUG_SHADING && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&P.stats[S_SHADOW], (unsigned long long)P.N);
}

// `*pixel += supersampling(...)` (src/renderer.rs:37,49-59), pass by pass in order
__global__ void __launch_bounds__(256) k_accumulate(RParams P) {
    for (uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x; pix < P.npix; pix += gridDim.x * blockDim.x) {
        D3 px = d3(P.accum[3 * (size_t)pix], P.accum[3 * (size_t)pix + 1], P.accum[3 * (size_t)pix + 2]);
        for (uint32_t pass = 0; pass < P.batch; pass++) {
            D3 acc = splat(0.0);
            size_t base = ((size_t)pass * P.npix + pix) * P.spp;
            for (uint32_t s = 0; s < P.spp; s++) acc = acc + d3(P.L[0][base + s], P.L[1][base + s], P.L[2][base + s]);
            px = px + acc;
        }
        P.accum[3 * (size_t)pix] = px.x; P.accum[3 * (size_t)pix + 1] = px.y; P.accum[3 * (size_t)pix + 2] = px.z;
    }
}

// ------------------------------------------------------------------------------------ resolve (src/renderer.rs:64-90)
struct ResolveParams {
    const double* accum;  // full image, row order, rgb
    double* tmp0; double* tmp1;
    uint8_t* rgb8;
    uint32_t W, H;
    double scale;
    hnm_config cfg;
};
__global__ void k_tonemap_gamma(ResolveParams R) {
    size_t n = (size_t)R.W * R.H;
    double inv_gamma = 1.0 / R.cfg.gamma_factor;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        D3 hdr = d3(R.accum[3 * i], R.accum[3 * i + 1], R.accum[3 * i + 2]) * R.scale;
        D3 ldr = hdr;
        if (R.cfg.tone_mapping_mode == 1) {  // src/tonemap.rs:22-27
            D3 color = hdr * R.cfg.tone_exposure;
            double luminance = 0.22 * color.x + 0.707 * color.y + 0.071 * color.z;
            double white_point = R.cfg.tone_white_point * R.cfg.tone_exposure;
            ldr = saturate(color * (luminance / (white_point * white_point) + 1.0) / (luminance + 1.0));
        }
        R.tmp0[3 * i] = dm::pow(ldr.x, inv_gamma);  // src/color.rs:38-48
        R.tmp0[3 * i + 1] = dm::pow(ldr.y, inv_gamma);
        R.tmp0[3 * i + 2] = dm::pow(ldr.z, inv_gamma);
    }
}
HNM_D double gaussian(double x, double sigma) {  // src/filter.rs:13-15
    return dm::exp(-(x * x) / (2.0 * sigma * sigma)) / (2.0 * HNM_PI * sigma * sigma);
}
__global__ void k_bilateral(ResolveParams R, const double* src, double* dst) {  // src/filter.rs:32-58
    size_t n = (size_t)R.W * R.H;
    const uint32_t width = R.W, height = R.H;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t x = (uint32_t)i % width, y = (uint32_t)i / width;
        D3 pixel = d3(src[3 * i], src[3 * i + 1], src[3 * i + 2]);
        double current_sum = pixel.x + pixel.y + pixel.z;
        double sum_scale = 1.0 / 3.0;
        D3 filtered = splat(0.0);
        double w_p = 0.0;
        uint32_t diameter = R.cfg.bilateral_diameter, half = diameter / 2;
        for (uint32_t a = 0; a < diameter; a++) {
            for (uint32_t b = 0; b < diameter; b++) {
                uint32_t nx = clamp_u32(x - (half - a), 0u, width - 1u);   // wrapping u32, as the release build
                uint32_t ny = clamp_u32(y - (half - b), 0u, height - 1u);
                size_t j = (size_t)ny * width + nx;
                D3 nb = d3(src[3 * j], src[3 * j + 1], src[3 * j + 2]);
                double nsum = nb.x + nb.y + nb.z;
                double g_i = gaussian(sum_scale * (nsum - current_sum), R.cfg.bilateral_sigma_i);
                uint32_t dx = x - nx, dy = y - ny;
                double dist = __dsqrt_rn((double)(uint32_t)(dx * dx + dy * dy));  // src/filter.rs:7-11
                double g_s = gaussian(dist, R.cfg.bilateral_sigma_s);
                double w = g_i * g_s;
                filtered = filtered + nb * w;
                w_p += w;
            }
        }
        D3 out = filtered / w_p;
        dst[3 * i] = out.x; dst[3 * i + 1] = out.y; dst[3 * i + 2] = out.z;
    }
}
HNM_D uint8_t f64_as_u8(double v) {  // Rust `as u8`: saturating, NaN -> 0
    uint32_t u = __double2uint_rz(v);
    return (uint8_t)(u > 255u ? 255u : u);
}
__global__ void k_quantise(ResolveParams R, const double* src) {  // src/color.rs:10-16
    size_t n = (size_t)R.W * R.H * 3;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        R.rgb8[i] = f64_as_u8(255.0 * saturate(src[i]));
}
// gathered [rank][padded_rows][W][3] -> image row order
__global__ void k_deinterleave(const double* gathered, double* full, uint32_t W, uint32_t H, uint32_t padded_rows, uint32_t nranks, uint32_t tile_rows) {
    size_t n = (size_t)nranks * padded_rows * W;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t x = (uint32_t)(i % W);
        uint32_t lr = (uint32_t)((i / W) % padded_rows);
        uint32_t rank = (uint32_t)(i / ((size_t)W * padded_rows));
        uint32_t y = local_to_global_row_h(lr, rank, nranks, tile_rows);
        if (y >= H) continue;
        size_t o = ((size_t)y * W + x) * 3;
        full[o] = gathered[3 * i]; full[o + 1] = gathered[3 * i + 1]; full[o + 2] = gathered[3 * i + 2];
    }
}

// ------------------------------------------------------------------------------------ batch (unit parity) kernels
__global__ void __launch_bounds__(ISAAC_THREADS, 1) k_isaac_batch(const uint64_t* seeds, uint32_t n, uint32_t count, uint64_t* out) {
    extern __shared__ uint64_t smem_isaac[];
    uint64_t* mem = smem_isaac + threadIdx.x;
    for (uint32_t p = blockIdx.x * ISAAC_THREADS + threadIdx.x; p < n; p += gridDim.x * ISAAC_THREADS) {
        uint64_t* o = out + (size_t)p * count;
        isaac64_seed<ISAAC_THREADS>(mem, seeds[4 * p], seeds[4 * p + 1], seeds[4 * p + 2], seeds[4 * p + 3], [&](int i, uint64_t v) {
            int j = 255 - i;
            if (j < (int)count) o[j] = v;
        });
    }
}
__global__ void k_isaac_full_batch(const uint64_t* seeds, uint32_t n, uint32_t count, uint64_t* out) {
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        IsaacFull rng;
        rng.seed(seeds[4 * p], seeds[4 * p + 1], seeds[4 * p + 2], seeds[4 * p + 3]);
        for (uint32_t j = 0; j < count; j++) out[(size_t)p * count + j] = rng.next_u64();
    }
}
__global__ void k_intersect_batch(DScene sc, const hnm_ray* rays, uint32_t n, hnm_hit* hits) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        D3 o = d3(rays[i].origin), d = d3(rays[i].direction);
        Hit h = trace<false>(sc, o, d, nullptr);
        hnm_hit out;
        memset(&out, 0, sizeof(out));
        if (h.kind == LEAF_NONE) {
            // Intersection::empty() + skybox emission (src/scene.rs:26-39,398)
            D3 e = skybox_sample(sc, d);
            out.distance = sc.inf; out.albedo = hnm_vec3{1.0, 1.0, 1.0}; out.emission = hnm_vec3{e.x, e.y, e.z};
            out.roughness = 0.2; out.hit = 0; out.element = -1; out.face = -1; out.surface = HNM_SURFACE_DIFFUSE;
        } else {
            uint32_t el = h.kind == LEAF_TRI ? sc.tri_elem[h.id] : h.id;
            const DMaterial& dm_ = sc.materials[sc.elements[el].material];
            SurfacePoint sp = surface_point(sc, h, o, d, true);
            PointMaterial pm = resolve_material(sc, dm_, sp.u, sp.v);
            out.position = hnm_vec3{sp.position.x, sp.position.y, sp.position.z};
            out.normal = hnm_vec3{sp.normal.x, sp.normal.y, sp.normal.z};
            out.albedo = hnm_vec3{pm.albedo.x, pm.albedo.y, pm.albedo.z};
            out.emission = hnm_vec3{pm.emission.x, pm.emission.y, pm.emission.z};
            out.distance = h.t; out.u = sp.u; out.v = sp.v; out.roughness = pm.roughness; out.param = pm.param;
            out.hit = 1; out.element = sp.element; out.face = sp.face; out.surface = pm.surface;
        }
        hits[i] = out;
    }
}
__global__ void k_material_sample_batch(DScene sc, const double* in, uint32_t n, double* out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double* p = in + 14 * (size_t)i;
        PointMaterial m;
        m.surface = (int32_t)p[0]; m.param = p[1]; m.roughness = p[2];
        m.albedo = splat(1.0); m.emission = splat(0.0);
        Rand2 rnd{p[3], p[4]};
        double s, c;
        dm::sincos(HNM_PI2 * rnd.r0, s, c);
        SampleResult r;
        r.origin = splat(0.0); r.direction = splat(0.0); r.reflectance = 0.0;
        bool some = material_sample(sc, m, rnd, c, s, d3(p[5], p[6], p[7]), d3(p[8], p[9], p[10]), d3(p[11], p[12], p[13]), r);
        double* o = out + 8 * (size_t)i;
        o[0] = some ? 1.0 : 0.0;
        o[1] = some ? r.origin.x : 0.0; o[2] = some ? r.origin.y : 0.0; o[3] = some ? r.origin.z : 0.0;
        o[4] = some ? r.direction.x : 0.0; o[5] = some ? r.direction.y : 0.0; o[6] = some ? r.direction.z : 0.0;
        o[7] = some ? r.reflectance : 0.0;
    }
}
__global__ void k_material_bsdf_batch(const double* in, uint32_t n, double* out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double* p = in + 12 * (size_t)i;
        PointMaterial m;
        m.surface = (int32_t)p[0]; m.param = p[1]; m.roughness = p[2];
        m.albedo = splat(1.0); m.emission = splat(0.0);
        out[i] = bsdf(m, d3(p[3], p[4], p[5]), d3(p[6], p[7], p[8]), d3(p[9], p[10], p[11]));
    }
}
__global__ void k_math_batch(int fn, const double* x, const double* y, uint32_t n, double* out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double r;
        switch (fn) {
            case 0: r = dm::sin(x[i]); break;
            case 1: r = dm::cos(x[i]); break;
            case 2: r = dm::exp(x[i]); break;
            case 3: r = dm::pow(x[i], y[i]); break;
            default: r = dm::acos(x[i]); break;
        }
        out[i] = r;
    }
}

}  // namespace hnm

// ====================================================================================== host side of the ABI
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
    double *tmp0 = nullptr, *tmp1 = nullptr, *full = nullptr;
    uint8_t* rgb8 = nullptr;
    bool profiling = false, trace_stats = false;
    uint64_t launches = 0;
    KernelTimer timer;
    int sm_count = 148;
    cudaEvent_t marks[16] = {};
};

namespace {

template <typename F>
void launch_timed(hnm_renderer* r, const char* name, F&& f) {
    r->launches++;
    if (!r->profiling) { f(); return; }
    cudaEvent_t a = r->timer.get(), b = r->timer.get();
    cudaEventRecord(a, r->stream);
    f();
    cudaEventRecord(b, r->stream);
    r->timer.pending.push_back({name, a, b});
}

template <typename T>
int dev_alloc(hnm_renderer* r, T** out, size_t count) {
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(count * sizeof(T), 16));
    if (e != cudaSuccess) return set_error(HNM_ERR_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    r->allocs.push_back(p);
    *out = (T*)p;
    return 0;
}

int run_batch(hnm_renderer* r, uint32_t sampling_first, uint32_t batch) {
    RParams& P = r->P;
    P.batch = batch;
    P.sampling_first = sampling_first;
    P.N = batch * P.npix * P.spp;
    if (P.N == 0) return 0;
    cudaStream_t st = r->stream;
    HNM_CUDA(cudaMemsetAsync(P.counters, 0, sizeof(uint32_t) * (OVF_COUNTER + 4), st));
    const int grid = r->sm_count * 4;
    if (P.mode == HNM_MODE_PATHTRACING) {
        size_t smem = (size_t)ISAAC_THREADS * 256 * sizeof(uint64_t);
        launch_timed(r, "isaac_raygen", [&] { k_isaac_raygen<<<r->sm_count, ISAAC_THREADS, smem, st>>>(P); });
        launch_timed(r, "rng_overflow", [&] { k_rng_overflow<<<r->sm_count, 64, 0, st>>>(P); });
        int buf = 0;
        for (int b = 1; b < (int)P.sc.bounce_limit; b++) {
            if (r->trace_stats) {
                launch_timed(r, "extend", [&] { k_extend<true><<<grid, 256, 0, st>>>(P, b, buf, 1); });
                launch_timed(r, "shade_miss", [&] { k_shade_miss<<<grid, 256, 0, st>>>(P, b, buf); });
                launch_timed(r, "shade_delta", [&] { k_shade_surf<false, true><<<grid, 256, 0, st>>>(P, b, buf); });
                launch_timed(r, "shade_nee", [&] { k_shade_surf<true, true><<<grid, 256, 0, st>>>(P, b, buf); });
            } else {
                launch_timed(r, "extend", [&] { k_extend<false><<<grid, 256, 0, st>>>(P, b, buf, 1); });
                launch_timed(r, "shade_miss", [&] { k_shade_miss<<<grid, 256, 0, st>>>(P, b, buf); });
                launch_timed(r, "shade_delta", [&] { k_shade_surf<false, false><<<grid, 256, 0, st>>>(P, b, buf); });
                launch_timed(r, "shade_nee", [&] { k_shade_surf<true, false><<<grid, 256, 0, st>>>(P, b, buf); });
            }
            buf ^= 1;
        }
    } else {
        launch_timed(r, "raygen_debug", [&] { k_raygen_debug<<<grid, 256, 0, st>>>(P); });
        launch_timed(r, "extend", [&] { k_extend<false><<<grid, 256, 0, st>>>(P, 1, 0, 0); });
        launch_timed(r, "debug_shade", [&] { k_debug_shade<<<grid, 256, 0, st>>>(P); });
    }
    launch_timed(r, "accumulate", [&] { k_accumulate<<<grid, 256, 0, st>>>(P); });
    HNM_CUDA(cudaGetLastError());
    return 0;
}

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
    if (r->stream) cudaStreamSynchronize(r->stream);
    r->timer.collect();
    for (auto p : r->allocs) cudaFree(p);
    for (auto e : r->marks) if (e) cudaEventDestroy(e);
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
    hnm_renderer* r = new hnm_renderer();
    r->scene = scene;
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
    uint32_t ntiles = (height + sh.tile_rows - 1) / sh.tile_rows;
    uint32_t tiles_per_rank = (ntiles + sh.num_ranks - 1) / sh.num_ranks;
    r->padded_rows = tiles_per_rank * sh.tile_rows;  // equal on every rank (all-gather)
    uint32_t real_rows = 0;
    for (uint32_t lr = 0; lr < r->padded_rows; lr++)
        if (local_to_global_row_h(lr, sh.rank, sh.num_ranks, sh.tile_rows) < height) real_rows++;
    // rows that exist are a prefix of the local rows (tiles are assigned in increasing order)
    P.real_rows = real_rows;
    P.npix = real_rows * width;
    cudaDeviceProp prop;
    HNM_CUDA(cudaGetDeviceProperties(&prop, scene->device));
    r->sm_count = prop.multiProcessorCount;
    size_t per_pass = (size_t)P.npix * P.spp;
    if (mode != HNM_MODE_PATHTRACING) max_batch = 1;
    if (max_batch == 0) {
        // enough paths in flight to fill the machine through the thin late bounces, bounded memory
        size_t target = 8u << 20;
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
    size_t cap = r->cap;
    for (int b = 0; b < 2; b++) {
        for (int k = 0; k < 6; k++) if ((rc = dev_alloc(r, &P.ray[b][k], cap))) return bail(rc);
        for (int k = 0; k < 3; k++) if ((rc = dev_alloc(r, &P.thr[b][k], cap))) return bail(rc);
        if ((rc = dev_alloc(r, &P.pid[b], cap))) return bail(rc);
    }
    for (int k = 0; k < 3; k++) if ((rc = dev_alloc(r, &P.L[k], cap))) return bail(rc);
    if ((rc = dev_alloc(r, &P.cursor, cap))) return bail(rc);
    if (mode == HNM_MODE_PATHTRACING) { if ((rc = dev_alloc(r, &P.tail, cap * RNG_TAIL))) return bail(rc); }
    if ((rc = dev_alloc(r, &P.hit_t, cap))) return bail(rc);
    if ((rc = dev_alloc(r, &P.hit_u, cap))) return bail(rc);
    if ((rc = dev_alloc(r, &P.hit_v, cap))) return bail(rc);
    if ((rc = dev_alloc(r, &P.hit_id, cap))) return bail(rc);
    if ((rc = dev_alloc(r, &P.q_miss, cap))) return bail(rc);
    if ((rc = dev_alloc(r, &P.q_delta, cap))) return bail(rc);
    if ((rc = dev_alloc(r, &P.q_nee, cap))) return bail(rc);
    if ((rc = dev_alloc(r, &P.q_ovf, cap))) return bail(rc);
    if ((rc = dev_alloc(r, &P.counters, (size_t)OVF_COUNTER + 4))) return bail(rc);
    if ((rc = dev_alloc(r, &P.stats, (size_t)S_COUNT))) return bail(rc);
    size_t accum_n = (size_t)r->padded_rows * width * 3;
    if ((rc = dev_alloc(r, &P.accum, accum_n))) return bail(rc);
    ce = cudaMemsetAsync(P.accum, 0, accum_n * sizeof(double), r->stream);
    if (ce == cudaSuccess) ce = cudaMemsetAsync(P.stats, 0, S_COUNT * sizeof(unsigned long long), r->stream);
    if (ce == cudaSuccess && mode == HNM_MODE_PATHTRACING)
        ce = cudaFuncSetAttribute(k_isaac_raygen, cudaFuncAttributeMaxDynamicSharedMemorySize, ISAAC_THREADS * 256 * (int)sizeof(uint64_t));
    if (ce != cudaSuccess) { set_error(HNM_ERR_CUDA, std::string("renderer init: ") + cudaGetErrorString(ce)); return bail(HNM_ERR_CUDA); }
    *out = r;
    return 0;
}

int hnm_render_passes(hnm_renderer* r, uint32_t sampling_first, uint32_t count) {
    if (!r) return set_error(HNM_ERR_INVALID, "null renderer");
    if (r->P.mode != HNM_MODE_PATHTRACING && count > 1) count = 1;  // DebugRenderer::max_sampling() == 1
    HNM_CUDA(cudaSetDevice(r->scene->device));
    uint32_t done = 0;
    while (done < count) {
        uint32_t b = std::min(r->max_batch, count - done);
        int rc = run_batch(r, sampling_first + done, b);
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

int hnm_resolve(hnm_renderer* r, const void* accum_full_device, uint32_t sampling, uint8_t* rgb8) {
    if (!r || !rgb8) return set_error(HNM_ERR_INVALID, "null argument");
    if (sampling == 0) return set_error(HNM_ERR_STATE, "resolve with sampling == 0");
    if (!accum_full_device && r->P.nranks != 1) return set_error(HNM_ERR_STATE, "a sharded renderer needs the gathered full-image buffer");
    HNM_CUDA(cudaSetDevice(r->scene->device));
    size_t n = (size_t)r->P.W * r->P.H;
    int rc = 0;
    if (!r->tmp0) {
        if ((rc = dev_alloc(r, &r->tmp0, n * 3))) return rc;
        if ((rc = dev_alloc(r, &r->tmp1, n * 3))) return rc;
        if ((rc = dev_alloc(r, &r->rgb8, n * 3))) return rc;
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
    HNM_CUDA(cudaMemcpyAsync(rgb8, r->rgb8, n * 3, cudaMemcpyDeviceToHost, st));
    HNM_CUDA(cudaStreamSynchronize(st));
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
        auto& t = r->timer.totals[name];
        names[k] = r->timer.totals.find(name)->first.c_str();
        ms[k] = (float)t.first; launches[k] = t.second;
        k++;
    }
    *n = k;
    return 0;
}

// ---- batch entry points -----------------------------------------------------------------------
#define HNM_TMP_UPLOAD(ptr, host, bytes) \
    HNM_CUDA(cudaMalloc(&ptr, std::max<size_t>(bytes, 16))); \
    HNM_CUDA(cudaMemcpy(ptr, host, bytes, cudaMemcpyHostToDevice));

int hnm_intersect_batch(hnm_scene* scene, const hnm_ray* rays, uint32_t n, hnm_hit* hits) {
    if (!scene || !rays || !hits) return set_error(HNM_ERR_INVALID, "null argument");
    if (n == 0) return 0;
    HNM_CUDA(cudaSetDevice(scene->device));
    hnm_ray* dr = nullptr; hnm_hit* dh = nullptr;
    HNM_TMP_UPLOAD(dr, rays, (size_t)n * sizeof(hnm_ray));
    HNM_CUDA(cudaMalloc(&dh, (size_t)n * sizeof(hnm_hit)));
    k_intersect_batch<<<592, 256>>>(scene->d, dr, n, dh);
    HNM_CUDA(cudaGetLastError());
    HNM_CUDA(cudaMemcpy(hits, dh, (size_t)n * sizeof(hnm_hit), cudaMemcpyDeviceToHost));
    cudaFree(dr); cudaFree(dh);
    return 0;
}
int hnm_isaac64_batch(int device, const uint64_t* seeds, uint32_t n, uint32_t count, uint64_t* out) {
    if (!seeds || !out) return set_error(HNM_ERR_INVALID, "null argument");
    if (n == 0 || count == 0) return 0;
    HNM_CUDA(cudaSetDevice(device));
    uint64_t *ds = nullptr, *dout = nullptr;
    HNM_TMP_UPLOAD(ds, seeds, (size_t)n * 4 * sizeof(uint64_t));
    HNM_CUDA(cudaMalloc(&dout, (size_t)n * count * sizeof(uint64_t)));
    if (count <= HNM_RNG_TAIL) {
        size_t smem = (size_t)ISAAC_THREADS * 256 * sizeof(uint64_t);
        HNM_CUDA(cudaFuncSetAttribute(k_isaac_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_isaac_batch<<<148, ISAAC_THREADS, smem>>>(ds, n, count, dout);
    } else {
        k_isaac_full_batch<<<148, 64>>>(ds, n, count, dout);  // the exact slow path, with refill
    }
    HNM_CUDA(cudaGetLastError());
    HNM_CUDA(cudaMemcpy(out, dout, (size_t)n * count * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    cudaFree(ds); cudaFree(dout);
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
    double *di = nullptr, *dout = nullptr;
    HNM_TMP_UPLOAD(di, in, (size_t)n * 14 * sizeof(double));
    HNM_CUDA(cudaMalloc(&dout, (size_t)n * 8 * sizeof(double)));
    DScene sc;
    default_math_scene(sc);
    k_material_sample_batch<<<296, 256>>>(sc, di, n, dout);
    HNM_CUDA(cudaGetLastError());
    HNM_CUDA(cudaMemcpy(out, dout, (size_t)n * 8 * sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(di); cudaFree(dout);
    return 0;
}
int hnm_material_bsdf_batch(int device, const double* in, uint32_t n, double* out) {
    if (!in || !out) return set_error(HNM_ERR_INVALID, "null argument");
    if (n == 0) return 0;
    HNM_CUDA(cudaSetDevice(device));
    double *di = nullptr, *dout = nullptr;
    HNM_TMP_UPLOAD(di, in, (size_t)n * 12 * sizeof(double));
    HNM_CUDA(cudaMalloc(&dout, (size_t)n * sizeof(double)));
    k_material_bsdf_batch<<<296, 256>>>(di, n, dout);
    HNM_CUDA(cudaGetLastError());
    HNM_CUDA(cudaMemcpy(out, dout, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(di); cudaFree(dout);
    return 0;
}
int hnm_math_batch(int device, int fn, const double* x, const double* y, uint32_t n, double* out) {
    if (!x || !y || !out) return set_error(HNM_ERR_INVALID, "null argument");
    if (n == 0) return 0;
    HNM_CUDA(cudaSetDevice(device));
    double *dx = nullptr, *dy = nullptr, *dout = nullptr;
    HNM_TMP_UPLOAD(dx, x, (size_t)n * sizeof(double));
    HNM_TMP_UPLOAD(dy, y, (size_t)n * sizeof(double));
    HNM_CUDA(cudaMalloc(&dout, (size_t)n * sizeof(double)));
    k_math_batch<<<296, 256>>>(fn, dx, dy, n, dout);
    HNM_CUDA(cudaGetLastError());
    HNM_CUDA(cudaMemcpy(out, dout, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(dx); cudaFree(dy); cudaFree(dout);
    return 0;
}

}  // extern "C"
