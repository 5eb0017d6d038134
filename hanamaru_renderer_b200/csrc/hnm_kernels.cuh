// hnm_kernels.cuh -- the wavefront kernel set around k_trace / k_confirm (hnm_trace.cuh).
//
// One batch of passes of PathTracingRenderer (src/renderer.rs:148-203):
//   k_isaac_raygen   ISAAC-64 seeding per path (rand 0.4 StdRng; 2 KB of state per path in shared memory,
//                    112 paths per CTA) + thin-lens camera ray (src/camera.rs:66-96); runs on the renderer's RNG
//                    stream, one batch ahead of the kernels below (hanamaru_b200.cu: GenSet)
//   k_rng_overflow   exact slow path for the (rare) paths whose lens rejection loop outruns the stored
//                    tail of the random stream
//   k_batch_begin    hands the generated batch to the renderer's stream (ray count, path / fallback statistics)
//   per bounce b = 1 .. bounce_limit-1:
//     k_trace        f32 candidate search; job 0: camera-path rays of bounce b, job 1: the NEE shadow rays of bounce b-1
//     k_confirm      exact closest hit of the camera rays from their candidate lists, classified into the
//                    miss / delta-BSDF / NEE-BSDF queues
//     k_nee_resolve  (b-1) exact closest hit of the shadow rays, visibility test + light contribution
//                    (src/renderer.rs:282-291), radiance update
//     k_shade_miss   Skybox::sample (src/scene.rs:295-319), radiance update, path ends
//     k_shade_surf   material resolve, BSDF sample, throughput update, compaction into the next ray queue;
//                    Diffuse / GGX hits also emit one shadow ray per emitter (src/renderer.rs:269-281)
//   k_trace + k_nee_resolve for the last bounce's shadow rays
//   k_accumulate     per pixel: sum of the sub-pixel paths in the reference's order, += into the f64
//                    accumulation buffer (src/renderer.rs:37,56)
// Resolve (src/renderer.rs:64-90): k_tonemap_gamma -> k_bilateral -> k_quantise.
// Queue sizes live in device memory; a batch is enqueued without any host synchronisation.
#ifndef HNM_KERNELS_CUH
#define HNM_KERNELS_CUH

#include "hnm_device.cuh"
#include "hnm_trace.cuh"

namespace hnm {

constexpr int RNG_TAIL = HNM_RNG_TAIL;  // u64 outputs kept per path
// ISAAC-64 seeding keeps the 2 KB `mem[]` of every path in shared memory: 112 paths x 2 KB = 224 KB of the 227 KB a CTA
// may use, i.e. only 112 paths in flight per SM, each a ~12 k-instruction dependent chain.  With full warps that is
// 3.5 warps per SM -- less than one per scheduler, pure latency (ncu, round 1: 35 % issue utilisation).  The paths
// are therefore spread over MORE warps with only ISAAC_LANES active lanes each (16 -> 7 warps, ~2 per scheduler):
// the same 112 chains, twice the warps to interleave.  Issue slots were idle anyway.
#ifndef HNM_ISAAC_LANES
#define HNM_ISAAC_LANES 32
#endif
#ifndef HNM_ISAAC_PATHS
#define HNM_ISAAC_PATHS 112
#endif
constexpr int ISAAC_PATHS = HNM_ISAAC_PATHS;  // paths (columns of the shared-memory state) per CTA
constexpr int ISAAC_LANES = HNM_ISAAC_LANES;
constexpr int ISAAC_THREADS = (ISAAC_PATHS + ISAAC_LANES - 1) / ISAAC_LANES * 32;
static_assert(ISAAC_LANES >= 1 && ISAAC_LANES <= 32, "bad ISAAC lane split");
// path column of this thread inside the CTA, or -1 for a lane that idles
__device__ __forceinline__ int isaac_slot() {
    const int lane = threadIdx.x & 31;
    const int slot = (int)(threadIdx.x >> 5) * ISAAC_LANES + lane;
    return (lane < ISAAC_LANES && slot < ISAAC_PATHS) ? slot : -1;
}
constexpr int MAX_BOUNCE = 64;

// counters[]: per bounce b (1-origin) eight slots
enum { C_RAY = 0, C_MISS = 1, C_DELTA = 2, C_NEE = 3, C_EVENTS = 4, C_SHADOW = 5, C_WORK = 6,
       // dynamic work counters of the kernels after k_trace (zero at batch start): see fetch_warp / fetch_cta
       C_W_CONFIRM = 7, C_W_MISS = 8, C_W_DELTA = 9, C_W_NEE = 10, C_W_NEER = 11, C_STRIDE = 16 };
constexpr int NUM_COUNTERS = (MAX_BOUNCE + 2) * C_STRIDE + 8;
// stats[] (u64)
enum { S_PATHS = 0, S_SEGMENTS = 1, S_SHADOW = 2, S_RNG_FALLBACK = 3, S_NODES = 4, S_PRIMS = 5, S_OVERFLOW = 6, S_COUNT = 8 };

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
    // camera-path rays, double buffered by bounce parity (queue order)
    // (`rin` = the queue this bounce reads, `rout` = the queue it writes; the host swaps them per launch so that no
    // kernel indexes a parameter array at run time)
    double* rin[6]; double* tin[3]; uint32_t* pin;
    double* rout[6]; double* tout[3]; uint32_t* pout;
    // per path (slot order)
    double* L[3];
    uint8_t* cursor;
    uint64_t* tail;       // [RNG_TAIL][cap]
    // hits of the current bounce (queue order)
    double* hit_t; double* hit_u; double* hit_v; uint2* hit_id;
    uint32_t* q_miss; uint32_t* q_delta; uint32_t* q_nee;
    // exact-slow-path list of the generation set that L / cursor / tail belong to (see GenSet, hanamaru_b200.cu)
    uint32_t* q_ovf; uint32_t* ovf_counter;
    // sliced generation (k_isaac_raygen_tm): next path to hand out (per generation set), stop level (per renderer), this launch's level
    uint32_t* gen_next; const uint32_t* gen_stop; uint32_t gen_epoch;
    // NEE events of the current bounce and their shadow rays (num_emissions per event, event-major)
    double* ev_thr[3]; double* ev_albedo[3]; double* ev_emission[3]; uint32_t* ev_pid;
    double* sray[6]; double* s_pos[3]; double* s_bsdf; double* s_g;
    float* s_tmax;        // distance from the shading point to the light sample: the shadow query is bounded (k_trace)
    uint32_t* counters;
    unsigned long long* stats;
    double* accum;        // [padded_rows * W * 3]
    unsigned long long* dbg;  // diagnostics (HNM_WID_STATS=1): masks of the hardware warp slots each kernel's warps ran in
};
// hardware warp slot of the calling warp inside its SM (the issue arbiter prefers high slots, B300_MICROARCH.md)
HNM_D void note_warp_slot(unsigned long long* dbg, int which) {
    if (dbg && (threadIdx.x & 31) == 0) {
        unsigned w;
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(w));
        atomicOr(&dbg[which], 1ull << (w & 63u));
    }
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
    c.y = local_to_global_row_h(lr, P.rank, P.nranks, P.tile_rows);
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

// `mem` is this thread's column of a [256][T] u64 array (shared memory: conflict-free for any per-lane
// index because the bank depends only on the lane).  Outputs rsl[i] are handed to `sink`.
// init(true) of rand 0.4.3: the two mixing passes over rsl = [s0 s1 s2 s3 0 0 ...] and then over mem itself
template <int T>
__device__ __forceinline__ void isaac64_init(uint64_t* mem, uint64_t s0, uint64_t s1, uint64_t s2, uint64_t s3) {
#define MEM(i) mem[(i) * T]
    uint64_t a, b, c, d, e, f, g, h;
    a = b = c = d = e = f = g = h = 0x9e3779b97f4a7c13ull;
#pragma unroll
    for (int i = 0; i < 4; i++) { ISAAC_MIX(a, b, c, d, e, f, g, h) }  // constants: folded at compile time
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
#undef MEM
}

// Shared-memory accesses of the round are volatile PTX with compile-time offsets from a row register: their ORDER in the
// instruction stream is the order written here (ptxas would otherwise sink the static-index loads next to their uses,
// which a single in-order warp then waits for).
// ~(a ^ b) as one LOP3 per half (ptxas emitted the xor and the not separately)
HNM_D uint64_t xnor64(uint64_t a, uint64_t b) {
    uint32_t lo, hi;
    asm("lop3.b32 %0, %1, %2, 0, 0xC3;" : "=r"(lo) : "r"((uint32_t)a), "r"((uint32_t)b));
    asm("lop3.b32 %0, %1, %2, 0, 0xC3;" : "=r"(hi) : "r"((uint32_t)(a >> 32)), "r"((uint32_t)(b >> 32)));
    return (uint64_t)lo | ((uint64_t)hi << 32);
}
template <int OFF>
HNM_D uint64_t lds64v(uint32_t saddr) {
    uint64_t v;
    asm volatile("ld.volatile.shared.u64 %0, [%1+%2];" : "=l"(v) : "r"(saddr), "n"(OFF) : "memory");
    return v;
}
template <int OFF>
HNM_D void sts64v(uint32_t saddr, uint64_t v) {
    asm volatile("st.volatile.shared.u64 [%0+%1], %2;" :: "r"(saddr), "n"(OFF), "l"(v) : "memory");
}
// The first isaac64() round after init: the 256 data-dependent steps over `mem` (this thread's column of [256][T] u64).
//   One step:  x = mem[i];  aa = mix(aa) + mem[i ^ 128];  y = mem[ind(x)] + aa + bb;  mem[i] = y;
//              bb = mem[ind(y >> 8)] + x;  rsl[i] = bb.
// One warp per scheduler issues in order, so the step time is the dependent chain plus every stall the in-order stream
// exposes.  The chain is  y -> address of mem[ind(y >> 8)] -> shared-memory load -> next y  and nothing else is on it:
//   * bb only ever enters the next y, so the step carries the raw load `ldy` and y = p + ldy + s is ONE three-input
//     addition behind it, with s = aa + x(previous) computed while the load is in flight;
//   * the other data-dependent load, p = mem[ind(x)] of the NEXT step, is issued right behind this step's store
//     (program order makes it see y if it aliases); its address was computed a step earlier from x, which is a
//     static-index load of a word no earlier step has written and is fetched two steps ahead;
//   * mem[i ^ 128] is fetched two steps ahead, so that aa and s of the NEXT step are computed in the shadow of this
//     step's loads.
// Round 1 loaded mem[ind(x)] BEFORE the store and patched the possible alias by forwarding y (a compare and two selects on
// the chain): 28 instructions and 94 cycles per step; this form has 21 instructions.
template <int T, int KEEP, typename Sink>
__device__ __forceinline__ void isaac64_round(uint64_t* mem, Sink sink) {
    static_assert(8 * T < 2048, "the umulhi form needs (8 * T) << 21 to fit 32 bits");
    static_assert(KEEP >= 8 && KEEP % 4 == 0 && KEEP <= 120, "KEEP");
    constexpr int ROW = 8 * T;  // bytes between mem[i] and mem[i + 1]
    const uint32_t col = (uint32_t)__cvta_generic_to_shared(mem);
    // byte address of mem[(x >> 3) & 255] is col + (x & 0x7f8) * T, that of mem[(y >> 11) & 255] the high word of
    // (y & 0x7f800) * (ROW << 21) + (col << 32); the addend is opaque so that ptxas keeps the register pair alive
    uint64_t cbase = (uint64_t)col << 32;
    asm("" : "+l"(cbase));
#define ADDRX(x) (col + ((uint32_t)(x) & 0x7f8u) * (uint32_t)T)
#define ADDRY(y) ((uint32_t)(((uint64_t)((uint32_t)(y) & 0x7f800u) * (uint64_t)((uint32_t)ROW << 21) + cbase) >> 32))
    // Software pipeline.  At the top of step i everything y_i needs except the two data-dependent loads is in registers:
    //   s = aa_i + x_(i-1), ax1 = address of mem[ind(x_(i+1))], x = x_i, x1, x2, m2 = mem[(i + 1) ^ 128].
    // The step issues, in this order: y, store, p_(i+1), ldy_i, then the static-index fetches x_(i+3) and mem[(i+2)^128],
    // and fills the shadow of the loads with aa_(i+1), s_(i+1) and the address for x_(i+2).
    // isaac64() after init: a = b = 0, c = 1  ->  aa = 0, bb = 1: ldy = 1 with x_(-1) = 0.
    uint64_t ldy = 1;
    uint64_t x = lds64v<0>(col), x1 = lds64v<ROW>(col), x2 = lds64v<2 * ROW>(col);
    uint64_t aa = 0xFFFFFFFFFFFFFFFFull + lds64v<128 * ROW>(col);  // aa_0 = ~(0 ^ (0 << 21)) + mem[128]
    uint64_t m2 = lds64v<129 * ROW>(col);
    uint64_t p = lds64v<0>(ADDRX(x));
    uint32_t ax1 = ADDRX(x1);
    uint64_t s = aa;
    uint64_t xs = 0;  // x of the previous step: rsl[i - 1] = ldy + xs is handed out one step late, when ldy has arrived
    // D = i - base, DX = min(i + 3, 255) - base, DM = ((i + 2) ^ 128) - base, mixexpr = the mix of step i + 1
#define ISAAC_STEP(mixexpr, row, base, D, DX, DM, SINK)                                    \
    {                                                                                     \
        const uint64_t y = p + ldy + s;                                                   \
        SINK((base) + (D) - 1, ldy + xs);                                                 \
        const uint32_t ay = ADDRY(y);                                                     \
        sts64v<(D) * ROW>(row, y);                                                        \
        p = lds64v<0>(ax1);                                                               \
        ldy = lds64v<0>(ay);                                                              \
        const uint64_t x3 = lds64v<(DX) * ROW>(row);                                      \
        const uint64_t m2n = lds64v<(DM) * ROW>(row);                                     \
        aa = (mixexpr) + m2;                                                              \
        s = aa + x;                                                                       \
        asm("" : "+l"(s)); /* y stays ONE three-input addition behind the loads */        \
        ax1 = ADDRX(x2);                                                                  \
        xs = x; x = x1; x1 = x2; x2 = x3; m2 = m2n;                                       \
    }
#define ISAAC_4STEPS(row, base, HALF, SINK)                                                                 \
    ISAAC_STEP(aa ^ (aa >> 5), row, base, 0, 3, 2 + (HALF), SINK)                                           \
    ISAAC_STEP(aa ^ (aa << 12), row, base, 1, 4, 3 + (HALF), SINK)                                          \
    ISAAC_STEP(aa ^ (aa >> 33), row, base, 2, 5, 4 + (HALF), SINK)                                          \
    ISAAC_STEP(xnor64(aa, aa << 21), row, base, 3, 6, 5 + (HALF), SINK)
#define ISAAC_NOSINK(i, v)
    // only the last KEEP outputs (rsl[256-KEEP .. 255], the first KEEP words of the stream) are handed to `sink`
#define ISAAC_SINK(i, v) if ((i) >= 256 - KEEP) sink(i, v)
    uint32_t row = col;
#pragma unroll 1
    for (int base = 0; base < 124; base += 4, row += 4 * ROW) { ISAAC_4STEPS(row, base, 128, ISAAC_NOSINK) }
    // steps 124..127: (i + 2) ^ 128 wraps to mem[0], mem[1]
    ISAAC_STEP(aa ^ (aa >> 5), row, 124, 0, 3, 130, ISAAC_NOSINK)
    ISAAC_STEP(aa ^ (aa << 12), row, 124, 1, 4, 131, ISAAC_NOSINK)
    ISAAC_STEP(aa ^ (aa >> 33), row, 124, 2, 5, -124, ISAAC_NOSINK)
    ISAAC_STEP(xnor64(aa, aa << 21), row, 124, 3, 6, -123, ISAAC_NOSINK)
    row += 4 * ROW;
#pragma unroll 1
    for (int base = 128; base < 256 - KEEP; base += 4, row += 4 * ROW) { ISAAC_4STEPS(row, base, -128, ISAAC_NOSINK) }
#pragma unroll 1
    for (int base = 256 - KEEP; base < 252; base += 4, row += 4 * ROW) { ISAAC_4STEPS(row, base, -128, ISAAC_SINK) }
    // steps 252..255: nothing is left to prefetch (dummy loads of valid words)
    ISAAC_STEP(aa ^ (aa >> 5), row, 252, 0, 3, -126, ISAAC_SINK)
    ISAAC_STEP(aa ^ (aa << 12), row, 252, 1, 3, -125, ISAAC_SINK)
    ISAAC_STEP(aa ^ (aa >> 33), row, 252, 2, 3, -125, ISAAC_SINK)
    ISAAC_STEP(xnor64(aa, aa << 21), row, 252, 3, 3, -125, ISAAC_SINK)
    sink(255, ldy + xs);
#undef ISAAC_SINK
#undef ISAAC_NOSINK
#undef ISAAC_4STEPS
#undef ISAAC_STEP
#undef ADDRX
#undef ADDRY
}
template <int T, int KEEP, typename Sink>
__device__ __forceinline__ void isaac64_seed(uint64_t* mem, uint64_t s0, uint64_t s1, uint64_t s2, uint64_t s3, Sink sink) {
    isaac64_init<T>(mem, s0, s1, s2, s3);
    isaac64_round<T, KEEP>(mem, sink);
}

// ------------------------------------------------------------------------------------ ISAAC-64, TMEM-pipelined
// The seeding of one path is init (64 mixes: a sequential, ALU-bound chain that touches mem[] only in order) followed by
// one round (256 steps, each waiting for a data-dependent shared-memory load).  Shared memory holds 112 states, i.e. one
// warp per scheduler, and a single warp cannot overlap its own two phases.  Blackwell's tensor memory is a second 256 KB
// of on-chip storage per SM -- 128 lanes x 512 columns x 32 bit = exactly one 2 KB state per lane -- that tcgen05.st / .ld
// address with a warp-uniform column, which is all init needs.  So the CTA runs TWO warps per scheduler:
//   producer warp (4 + w): init of path k+1 with mem[] in its TMEM lanes (pass 1 stores, pass 2 loads and stores in
//                          place), then -- once the consumer is done with path k -- copies the state to shared memory;
//   consumer warp (w):     the round of path k in shared memory, the kept outputs, the lens sample and the camera ray.
// The pair meets at two named barriers per path.  The hardware interleaves the ALU-bound chain of one warp with the
// load-latency-bound chain of the other; the per-path time drops from init + round to about max(init + copy, round).
#ifndef HNM_TM_DIAG
#define HNM_TM_DIAG 0  /* timing diagnostics (wrong results): 1 no round, 2 no init and no copy, 3 no copy */
#endif
constexpr int ISAAC_TM_LANES = 28;                       // active consumer lanes per warp: 4 x 28 = 112 columns
constexpr int ISAAC_TM_THREADS = 256;
static_assert(ISAAC_PATHS == 4 * ISAAC_TM_LANES, "the TMEM pipeline is laid out for 4 consumer warps x 28 lanes");

HNM_D void tm_st16(uint32_t taddr, uint64_t a, uint64_t b, uint64_t c, uint64_t d, uint64_t e, uint64_t f, uint64_t g, uint64_t h) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 :: "r"(taddr),
                    "r"((uint32_t)a), "r"((uint32_t)(a >> 32)), "r"((uint32_t)b), "r"((uint32_t)(b >> 32)),
                    "r"((uint32_t)c), "r"((uint32_t)(c >> 32)), "r"((uint32_t)d), "r"((uint32_t)(d >> 32)),
                    "r"((uint32_t)e), "r"((uint32_t)(e >> 32)), "r"((uint32_t)f), "r"((uint32_t)(f >> 32)),
                    "r"((uint32_t)g), "r"((uint32_t)(g >> 32)), "r"((uint32_t)h), "r"((uint32_t)(h >> 32))
                 : "memory");
}
HNM_D void tm_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr)
                 : "memory");
}
HNM_D void tm_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr)
                 : "memory");
}
HNM_D void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
HNM_D void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
HNM_D void named_barrier(uint32_t id, uint32_t threads) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(threads) : "memory"); }
HNM_D uint64_t u64_of(uint32_t lo, uint32_t hi) { return (uint64_t)lo | ((uint64_t)hi << 32); }

// TMEM columns [c0, c1) of this warp's lanes -> words [c0 / 2, c1 / 2) of shared-memory column `mem` (stride T words).  Loads of
// 32 columns, the next one in flight while the 16 words of the current one are stored (with one x16 load at a time the
// loop was bound by the TMEM load latency: 0.93 -> 0.5 ms of the 7 ms kernel).  All 32 lanes must call.
#ifndef HNM_TM_COPY_X32
#define HNM_TM_COPY_X32 0
#endif
template <int T>
HNM_D void tm_copy_to_column(uint32_t tbase, uint64_t* mem, bool active, uint32_t c0, uint32_t c1) {
    if (c0 >= c1) return;
#if HNM_TM_COPY_X32
    uint32_t c[32];
    tm_ld32(tbase + c0, c);
#pragma unroll 1
    for (uint32_t col = c0; col < c1; col += 32) {
        tm_wait_ld();
        uint64_t w[16];
#pragma unroll
        for (int j = 0; j < 16; j++) w[j] = u64_of(c[2 * j], c[2 * j + 1]);
        tm_ld32(tbase + (col + 32 < c1 ? col + 32 : c0), c);  // (the last iteration re-reads the first block: unused)
        if (active) {
            uint64_t* o = mem + (size_t)(col >> 1) * T;
#pragma unroll
            for (int j = 0; j < 16; j++) o[j * T] = w[j];
        }
    }
    tm_wait_ld();
#else
    // two 16-column loads in flight, stored straight from their registers: the same one wait per 32 columns as the x32
    // form, without its 16-word staging copy (87 -> 64 registers for the whole kernel: the shade kernels beside a
    // generation slice get a third CTA per SM)
    uint32_t a[16], b[16];
    tm_ld16(tbase + c0, a);
    tm_ld16(tbase + c0 + 16, b);
#pragma unroll 1
    for (uint32_t col = c0; col < c1; col += 32) {
        tm_wait_ld();
        uint64_t* o = mem + (size_t)(col >> 1) * T;
        if (active) {
#pragma unroll
            for (int j = 0; j < 8; j++) o[j * T] = u64_of(a[2 * j], a[2 * j + 1]);
        }
        tm_ld16(tbase + (col + 32 < c1 ? col + 32 : c0), a);
        if (active) {
#pragma unroll
            for (int j = 0; j < 8; j++) o[(8 + j) * T] = u64_of(b[2 * j], b[2 * j + 1]);
        }
        tm_ld16(tbase + (col + 48 < c1 ? col + 48 : c0), b);
    }
    tm_wait_ld();
#endif
}
// Work is handed out per warp PAIR, 28 consecutive paths at a time: fetch() is called by the producer warp (every lane gets the
// same answer) and returns the first path of the next group or ISAAC_NONE; seed(p, s0..s3) is called by the producer lanes
// for path p = group + lane; sink(p, i, v) and done(p) by the consumer lanes (`done` runs while the state of the next group is
// copied in).  Every thread of the 256-thread CTA must call this.
constexpr uint32_t ISAAC_NONE = 0xFFFFFFFFu;
template <int KEEP, typename FetchFn, typename SeedFn, typename SinkFn, typename DoneFn>
__device__ __forceinline__ void isaac64_tmem_pipeline(uint64_t* smem, FetchFn fetch, SeedFn seed, SinkFn sink, DoneFn done) {
    constexpr int T = ISAAC_PATHS;
    __shared__ uint32_t s_tmem_base;
    __shared__ uint32_t s_group[4];  // the group whose state the producer has just copied into the pair's columns
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t pair = warp & 3u;
    const uint32_t bar_free = 1 + 2 * pair, bar_ready = 2 + 2 * pair;  // named barriers of the pair (64 threads)
    uint64_t* const mem = smem + pair * ISAAC_TM_LANES + (lane < (uint32_t)ISAAC_TM_LANES ? lane : 0u);
    if (warp < 4) {
        // ---------------- consumer: rounds in shared memory
        if (lane >= (uint32_t)ISAAC_TM_LANES) return;  // (a warp's barrier arrival does not depend on its exited lanes)
        constexpr unsigned CMASK = (1u << ISAAC_TM_LANES) - 1u;
        uint32_t prev = ISAAC_NONE;
        for (;;) {
            __syncwarp(CMASK);
            named_barrier(bar_free, 64);   // these columns are free: the producer may copy the next state in
#if HNM_TM_DIAG != 4
            if (prev != ISAAC_NONE) done(prev + lane);
#endif
            __syncwarp(CMASK);
            named_barrier(bar_ready, 64);  // the next state is in shared memory (or there is none)
            const uint32_t cur = *(volatile uint32_t*)&s_group[pair];
            if (cur == ISAAC_NONE) break;
#if HNM_TM_DIAG != 1
            const uint32_t path = cur + lane;
            isaac64_round<T, KEEP>(mem, [&](int i, uint64_t v) { sink(path, i, v); });
#endif
            prev = cur;
        }
        return;
    }
    // ---------------- producer: init in tensor memory (all 32 lanes execute the tcgen05 instructions; lanes 28-31 idle along)
    if (warp == 4) {
        uint32_t dst = (uint32_t)__cvta_generic_to_shared(&s_tmem_base);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(dst), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    named_barrier(15, 128);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = s_tmem_base + ((pair * 32u) << 16);  // lanes 32 * (warp % 4) .. + 31 belong to this warp
#ifdef HNM_TM_STAGGER
    {   // experiment: start the four pairs a quarter of an iteration apart
        const long long t0 = clock64();
        while (clock64() - t0 < (long long)pair * HNM_TM_STAGGER) {}
    }
#endif
    const bool active = lane < (uint32_t)ISAAC_TM_LANES;
    for (;;) {
        const uint32_t group = fetch();
        if (group == ISAAC_NONE) {
            named_barrier(bar_free, 64);
            if (lane == 0) *(volatile uint32_t*)&s_group[pair] = ISAAC_NONE;
            named_barrier(bar_ready, 64);
            break;
        }
        uint32_t m[16];
#if HNM_TM_DIAG != 2
        uint64_t s0, s1, s2, s3;
        seed(group + (active ? lane : 0u), s0, s1, s2, s3);
        uint64_t a, b, c, d, e, f, g, h;
        a = b = c = d = e = f = g = h = 0x9e3779b97f4a7c13ull;
#pragma unroll
        for (int i = 0; i < 4; i++) { ISAAC_MIX(a, b, c, d, e, f, g, h) }
        a += s0; b += s1; c += s2; d += s3;
#pragma unroll 1
        for (uint32_t col = 0; col < 512; col += 16) {
            ISAAC_MIX(a, b, c, d, e, f, g, h)
            tm_st16(tbase + col, a, b, c, d, e, f, g, h);
        }
        tm_wait_st();
        tm_ld16(tbase, m);
#pragma unroll 1
        for (uint32_t col = 0; col < 512; col += 16) {
            tm_wait_ld();
            a += u64_of(m[0], m[1]); b += u64_of(m[2], m[3]); c += u64_of(m[4], m[5]); d += u64_of(m[6], m[7]);
            e += u64_of(m[8], m[9]); f += u64_of(m[10], m[11]); g += u64_of(m[12], m[13]); h += u64_of(m[14], m[15]);
            tm_ld16(tbase + ((col + 16) & 511u), m);  // next block (the last iteration re-reads block 0: unused)
            ISAAC_MIX(a, b, c, d, e, f, g, h)
            tm_st16(tbase + col, a, b, c, d, e, f, g, h);
        }
        tm_wait_ld();
        tm_wait_st();
#endif
        named_barrier(bar_free, 64);
#if HNM_TM_DIAG != 2 && HNM_TM_DIAG != 3
        // copy the finished state into the consumer's columns.  (Letting the consumer warp copy a share itself -- its lanes
        // 28-31 kept alive for the warp-wide tcgen05.ld -- was measured slower for every split: 6.56 -> 8.1-8.6 ms.)
        tm_copy_to_column<T>(tbase, mem, active, 0, 512);
#endif
        if (lane == 0) *(volatile uint32_t*)&s_group[pair] = group;
        named_barrier(bar_ready, 64);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    named_barrier(14, 128);  // the producers only (no other warp touches tensor memory)
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(s_tmem_base), "r"(512u) : "memory");
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

HNM_D void store_ray(const RParams& P, uint32_t q, D3 o, D3 d, D3 t, uint32_t pid) {
    P.rout[0][q] = o.x; P.rout[1][q] = o.y; P.rout[2][q] = o.z;
    P.rout[3][q] = d.x; P.rout[4][q] = d.y; P.rout[5][q] = d.z;
    P.tout[0][q] = t.x; P.tout[1][q] = t.y; P.tout[2][q] = t.z;
    P.pout[q] = pid;
}

#ifndef HNM_ISAAC_MIN_BLOCKS
#define HNM_ISAAC_MIN_BLOCKS 1  /* 8 caps the kernel at 64 registers: one of its CTAs then fits next to 7 k_trace CTAs */
#endif
__global__ void __launch_bounds__(ISAAC_THREADS, HNM_ISAAC_MIN_BLOCKS) k_isaac_raygen(RParams P) {
    extern __shared__ uint64_t smem_isaac[];
    const int slot = isaac_slot();
    if (slot < 0) return;  // no CTA-wide synchronisation below
    note_warp_slot(P.dbg, 0);
    uint64_t* mem = smem_isaac + slot;
    const uint32_t N = P.N, cap = P.cap;
    for (uint32_t p = blockIdx.x * ISAAC_PATHS + slot; p < N; p += gridDim.x * ISAAC_PATHS) {
        PathCoord c = path_coord(P, p);
        uint64_t s0, s1, s2, s3;
        path_seed(c, P.sampling_first + c.pass, s0, s1, s2, s3);
        uint64_t* tail = P.tail + p;
        // outputs are consumed from rsl[255] downwards: word j of the stream = rsl[255 - j]
        isaac64_seed<ISAAC_PATHS, RNG_TAIL>(mem, s0, s1, s2, s3, [&](int i, uint64_t v) { tail[(size_t)(255 - i) * cap] = v; });
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
            // the stored tail is too short for this path: exact slow path (k_rng_overflow fills the slot)
            uint32_t slot = atomicAdd(P.ovf_counter, 1u);
            P.q_ovf[slot] = p;
            continue;
        }
        D3 o, d;
        lens_ray(P.cam, c.ncx, c.ncy, sqx, sqy, o, d);
        store_ray(P, p, o, d, splat(1.0), p);
        P.cursor[p] = (uint8_t)cur;
    }
}

// The same generation with the TMEM pipeline above (default; HNM_ISAAC_TMEM=0 selects k_isaac_raygen for the A/B).
#ifndef HNM_ISAAC_TM_MIN_BLOCKS
#define HNM_ISAAC_TM_MIN_BLOCKS 1
#endif
__global__ void __launch_bounds__(ISAAC_TM_THREADS, HNM_ISAAC_TM_MIN_BLOCKS) k_isaac_raygen_tm(RParams P) {
    extern __shared__ uint64_t smem_isaac[];
    const uint32_t N = P.N, cap = P.cap;
    const uint32_t lane = threadIdx.x & 31;
    if (threadIdx.x < 128) note_warp_slot(P.dbg, 0);
    isaac64_tmem_pipeline<RNG_TAIL>(
        smem_isaac,
        [&]() -> uint32_t {
            // The next 28 paths of the generation set, from a counter that survives the launch: a launch may be told to stop
            // (gen_stop >= gen_epoch, raised by k_gen_stop on the renderer's stream) and the next launch carries on where it
            // left off -- hanamaru_b200.cu runs the generation in slices beside the shade kernels of every bounce.
            uint32_t g = ISAAC_NONE;
            if (lane == 0) {
                const bool stop = *(volatile const uint32_t*)P.gen_stop >= P.gen_epoch;
                if (!stop) {
                    g = atomicAdd(P.gen_next, (uint32_t)ISAAC_TM_LANES);
                    if (g >= N) g = ISAAC_NONE;
                }
            }
            return __shfl_sync(0xFFFFFFFFu, g, 0);
        },
        [&](uint32_t p, uint64_t& s0, uint64_t& s1, uint64_t& s2, uint64_t& s3) {
            p = p < N ? p : N - 1;  // a column past the end idles along on the last path
            PathCoord c = path_coord(P, p);
            path_seed(c, P.sampling_first + c.pass, s0, s1, s2, s3);
        },
        [&](uint32_t p, int i, uint64_t v) {
            // outputs are consumed from rsl[255] downwards: word j of the stream = rsl[255 - j]
            if (p < N) P.tail[(size_t)(255 - i) * cap + p] = v;
        },
        [&](uint32_t p) {
            if (p >= N) return;
            PathCoord c = path_coord(P, p);
            const uint64_t* tail = P.tail + p;
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
                // the stored tail is too short for this path: exact slow path (k_rng_overflow fills the slot)
                uint32_t slot = atomicAdd(P.ovf_counter, 1u);
                P.q_ovf[slot] = p;
                return;
            }
            D3 o, d;
            lens_ray(P.cam, c.ncx, c.ncy, sqx, sqy, o, d);
            store_ray(P, p, o, d, splat(1.0), p);
            P.cursor[p] = (uint8_t)cur;
        });
}
// raises the stop level of the sliced generation (see k_isaac_raygen_tm)
__global__ void k_gen_stop(uint32_t* gen_stop, uint32_t epoch) { atomicMax(gen_stop, epoch); }

// First kernel of a path-tracing batch on the renderer's stream.  The generation kernels above may have run
// long before (on the RNG stream, overlapped with the previous batch), so they do not touch `counters` / `stats`.
__global__ void k_batch_begin(RParams P) {
    P.counters[1 * C_STRIDE + C_RAY] = P.N;
    atomicAdd(&P.stats[S_PATHS], (unsigned long long)P.N);
    atomicAdd(&P.stats[S_RNG_FALLBACK], (unsigned long long)*P.ovf_counter);
}

__global__ void k_rng_overflow(RParams P) {
    uint32_t n = *P.ovf_counter;
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
        store_ray(P, p, o, d, splat(1.0), p);
        P.cursor[p] = 0;
    }
}

// DebugRenderer: pinhole ray, no RNG (src/camera.rs:98-107, src/renderer.rs:117)
__global__ void k_raygen_debug(RParams P) {
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < P.N; p += gridDim.x * blockDim.x) {
        PathCoord c = path_coord(P, p);
        D3 o = d3(P.cam.eye);
        D3 d = normalize(c.ncx * d3(P.cam.plane_half_right) + c.ncy * d3(P.cam.plane_half_up) + P.cam.focus_distance * d3(P.cam.forward));
        store_ray(P, p, o, d, splat(1.0), p);
        P.L[0][p] = 0.0; P.L[1][p] = 0.0; P.L[2][p] = 0.0;
        P.cursor[p] = 0;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        P.counters[1 * C_STRIDE + C_RAY] = P.N;
        atomicAdd(&P.stats[S_PATHS], (unsigned long long)P.N);
    }
}

// Dynamic work distribution.  A static grid-stride loop gives every CTA of the grid the same share of the items; when
// fewer CTAs are resident than the grid has (the generation kernel of the NEXT batch holds an SM's shared memory and
// 8-9 K registers while these kernels run), the shares of the CTAs that did not fit run as a second, nearly empty wave.
// Fetching 32 (warp) or blockDim (CTA) consecutive items at a time from a counter lets whatever is resident drain the
// queue evenly.  Which warp handles which item never changes a result: every path's arithmetic is its own.
#ifndef HNM_DYN_MISS
#define HNM_DYN_MISS 1
#endif
#ifndef HNM_DYN_SURF
// k_shade_surf keeps its static grid-stride loop.  With dynamic fetching (either formulation below) the 64-register build
// of this kernel -- the one that spills to local memory -- gave wrong, run-to-run varying results on the one scene whose
// floor samples two image textures per hit (rtcamp5_pl), while compute-sanitizer racecheck / synccheck / initcheck were
// clean and the same source compiled for 80 or 128 registers was bit-exact again (the ptxas hazard).
// Measured gain of the dynamic variant where it was correct: 0.5 %.
#define HNM_DYN_SURF 0
#endif
#ifndef HNM_DYN_NEER
#define HNM_DYN_NEER 1
#endif
#ifndef HNM_DYN_CONFIRM
#define HNM_DYN_CONFIRM 1
#endif
HNM_D uint32_t fetch_warp(uint32_t* counter) {
    uint32_t base = 0;
    if ((threadIdx.x & 31) == 0) base = atomicAdd(counter, 32u);
    return __shfl_sync(0xFFFFFFFFu, base, 0);
}
HNM_D uint32_t fetch_cta(uint32_t* counter) {  // all threads of the CTA must call; the caller must place a barrier before the next call
    __shared__ uint32_t s_chunk;
    if (threadIdx.x == 0) s_chunk = atomicAdd(counter, blockDim.x);
    __syncthreads();
    return s_chunk;
}

// slot in a compacted queue for every lane with `pred`; one atomic per warp
HNM_D uint32_t queue_alloc(bool pred, uint32_t* counter) {
    unsigned mask = __ballot_sync(0xFFFFFFFFu, pred);
    if (mask == 0) return 0;
    int lane = threadIdx.x & 31;
    int leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(mask));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    return base + __popc(mask & ((1u << lane) - 1u));
}

// Slot in the next ray queue for every thread of the CTA with `pred`, grouped by `key` (0..7) inside the CTA's
// contiguous run of slots: the rays a CTA emits come from neighbouring paths (similar origins); ordering them by
// direction octant makes the warps of the next k_trace launch pick up rays that also descend the tree the same
// way.  One global atomic per CTA.  All threads of the CTA must call this (two barriers).
#ifndef HNM_SHADE_OCTANT_SORT
#define HNM_SHADE_OCTANT_SORT 1
#endif
HNM_D uint32_t queue_alloc_grouped(bool pred, uint32_t key, uint32_t* counter) {
    __shared__ uint32_t s_cnt[8], s_base[8], s_gbase;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < 8) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t k = pred ? key : 8u;
    const unsigned peers = __match_any_sync(0xFFFFFFFFu, k);
    const int leader = __ffs(peers) - 1;
    uint32_t woff = 0;
    if (pred && lane == leader) woff = atomicAdd(&s_cnt[k], (uint32_t)__popc(peers));
    woff = __shfl_sync(0xFFFFFFFFu, woff, leader);
    const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tot = 0;
        for (int j = 0; j < 8; j++) { s_base[j] = tot; tot += s_cnt[j]; }
        s_gbase = tot ? atomicAdd(counter, tot) : 0u;
    }
    __syncthreads();
    return pred ? s_gbase + s_base[k] + woff + rank : 0u;
}
// The same grouping inside ONE warp's run of slots: no shared memory, no CTA barrier (A/B: HNM_SHADE_OCTANT_SORT=2).
HNM_D uint32_t queue_alloc_warp_grouped(bool pred, uint32_t key, uint32_t* counter) {
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    const unsigned all = __ballot_sync(FULL, pred);
    if (all == 0) return 0;
    uint32_t before = 0, mine = 0;
#pragma unroll
    for (uint32_t k = 0; k < 8; k++) {
        const unsigned m = __ballot_sync(FULL, pred && key == k);
        if (k < key) before += __popc(m);
        if (k == key) mine = __popc(m & ((1u << lane) - 1u));
    }
    const int leader = __ffs(all) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(all));
    base = __shfl_sync(FULL, base, leader);
    return base + before + mine;
}
HNM_D uint32_t direction_octant(D3 d) { return (signbit_(d.x) ? 1u : 0u) | (signbit_(d.y) ? 2u : 0u) | (signbit_(d.z) ? 4u : 0u); }

HNM_D D3 load_ray_o(const RParams& P, uint32_t q) { return d3(P.rin[0][q], P.rin[1][q], P.rin[2][q]); }
HNM_D D3 load_ray_d(const RParams& P, uint32_t q) { return d3(P.rin[3][q], P.rin[4][q], P.rin[5][q]); }
HNM_D D3 load_thr(const RParams& P, uint32_t q) { return d3(P.tin[0][q], P.tin[1][q], P.tin[2][q]); }

// ------------------------------------------------------------------------------------ shade
HNM_D Rand2 bounce_random(const RParams& P, uint32_t pid, int bounce) {
    // `let random = rng.gen::<(f64, f64)>()` at the top of every bounce (src/renderer.rs:175)
    size_t w = (size_t)P.cursor[pid] + 2u * (uint32_t)(bounce - 1);
    Rand2 r;
    r.r0 = u64_to_f64(P.tail[w * P.cap + pid]);
    r.r1 = u64_to_f64(P.tail[(w + 1) * P.cap + pid]);
    return r;
}

#ifndef HNM_MISS_MIN_BLOCKS
#define HNM_MISS_MIN_BLOCKS 4  /* 64 registers: 0.68 ms where the unconstrained build varied 0.68-0.93 */
#endif
#ifndef HNM_NEER_MIN_BLOCKS
#define HNM_NEER_MIN_BLOCKS 4  /* 64 registers: 0.70 -> 0.60 ms */
#endif
template <bool FAST>
__global__ void __launch_bounds__(256, HNM_MISS_MIN_BLOCKS) k_shade_miss(RParams P, int bounce) {
    const uint32_t n = P.counters[bounce * C_STRIDE + C_MISS];
    uint32_t* const work = &P.counters[bounce * C_STRIDE + C_W_MISS];
#if HNM_DYN_MISS
    for (;;) {
        const uint32_t base = fetch_warp(work);
        if (base >= n) break;
        const uint32_t i = base + (threadIdx.x & 31);
        if (i >= n) continue;
#else
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#endif
        uint32_t q = P.q_miss[i];
        uint32_t pid = P.pin[q];
        D3 d = load_ray_d(P, q);
        D3 thr = load_thr(P, q);
        D3 emission = skybox_sample<FAST>(P.sc, d);  // src/scene.rs:398
        // accumulation += reflectance * emission (src/renderer.rs:196); the path ends (!hit, :199)
        D3 L = d3(P.L[0][pid], P.L[1][pid], P.L[2][pid]);
        L = L + thr * emission;
        P.L[0][pid] = L.x; P.L[1][pid] = L.y; P.L[2][pid] = L.z;
    }
}

// One surface interaction of PathTracingRenderer::calc_pixel (src/renderer.rs:176-199) for queue
// `cls` (C_DELTA: Specular / Refraction / GGXRefraction, C_NEE: Diffuse / GGX).  For the NEE class the
// light samples (Sphere::sample_on_surface, src/scene.rs:92-101) become shadow rays for the next k_trace
// and the radiance update moves to k_nee_resolve, which keeps the reference's order
// `accumulation += reflectance * nee` BEFORE `accumulation += reflectance * emission`.
#ifndef HNM_SHADE_MIN_BLOCKS
#define HNM_SHADE_MIN_BLOCKS 4  /* 64 registers (spills to local memory): measured best of 2 / 3 / 4 / 5 / 6 */
#endif
template <bool NEE, bool FAST>
#ifndef HNM_SURF_THREADS
#define HNM_SURF_THREADS 256  /* CTA size of k_shade_surf = the domain of the octant grouping (A/B: 512 with 2 CTAs/SM) */
#endif
__global__ void __launch_bounds__(HNM_SURF_THREADS, HNM_SHADE_MIN_BLOCKS * 256 / HNM_SURF_THREADS) k_shade_surf(RParams P, int bounce) {
    const int cls = NEE ? C_NEE : C_DELTA;
    note_warp_slot(P.dbg, 2);
    const uint32_t n = P.counters[bounce * C_STRIDE + cls];
    const uint32_t* queue = NEE ? P.q_nee : P.q_delta;
    const bool last_bounce = (uint32_t)bounce + 1 >= P.sc.bounce_limit;
    const uint32_t nl = P.sc.num_emissions;
    uint32_t shadow_rays = 0;
    uint32_t* const work = &P.counters[bounce * C_STRIDE + (NEE ? C_W_NEE : C_W_DELTA)];
#if HNM_DYN_SURF == 2
    for (;;) {
        // every warp fetches its own 32 items; the CTA leaves the loop together (barriers inside the body)
        const uint32_t wbase = fetch_warp(work);
        if (!__syncthreads_or(wbase < n ? 1 : 0)) break;
        const uint32_t i = wbase < n ? wbase + (threadIdx.x & 31) : 0xFFFFFFFFu;
#elif HNM_DYN_SURF
    for (;;) {
        __syncthreads();  // the previous iteration's readers of the chunk base are done
        const uint32_t chunk = fetch_cta(work);  // CTA-uniform: barriers inside the body
        if (chunk >= n) break;
        const uint32_t i = chunk + threadIdx.x;
#else
    const uint32_t n_round = (n + 255u) & ~255u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
#endif
        bool alive = false, event = false;
        D3 no = splat(0.0), nd = splat(0.0), nthr = splat(0.0), thr = splat(0.0), view = splat(0.0);
        SurfacePoint sp;
        PointMaterial pm;
        Rand2 random;
        double cos_phi = 1.0, sin_phi = 0.0;
        uint32_t pid = 0;
        if (i < n) {
            uint32_t q = queue[i];
            pid = P.pin[q];
            D3 o = load_ray_o(P, q), d = load_ray_d(P, q);
            thr = load_thr(P, q);
            Hit h;
            h.t = P.hit_t[q]; h.u = P.hit_u[q]; h.v = P.hit_v[q];
            uint2 hid = P.hit_id[q];
            h.kind = hid.x; h.id = hid.y;
            uint32_t el = h.kind == LEAF_TRI ? P.sc.tri_elem[h.id] : h.id;
            const DMaterial& dm_ = P.sc.materials[P.sc.elements[el].material];
            sp = surface_point<FAST>(P.sc, h, o, d, dm_.has_image != 0);
            pm = resolve_material<FAST>(P.sc, dm_, sp.u, sp.v);
            random = bounce_random(P, pid, bounce);
            if (pm.surface != HNM_SURFACE_SPECULAR && pm.surface != HNM_SURFACE_REFRACTION) {
                if (FAST && HNM_FAST_SINCOS) {
                    float sf, cf;
                    sincosf((float)(HNM_PI2 * random.r0), &sf, &cf);
                    sin_phi = sf; cos_phi = cf;
                } else {
                    dm::sincos(HNM_PI2 * random.r0, sin_phi, cos_phi);
                }
            }
            view = -d;
            SampleResult res;
            bool some = material_sample(P.sc, pm, random, cos_phi, sin_phi, sp.position, view, sp.normal, res);
            if (some) {
                nthr = thr * (pm.albedo * res.reflectance);          // src/renderer.rs:197
                alive = !all_zero(nthr) && !last_bounce;             // :199 and the loop bound :174
                no = res.origin; nd = res.direction;
                if (NEE) {
                    event = true;
                } else {
                    D3 L = d3(P.L[0][pid], P.L[1][pid], P.L[2][pid]);
                    L = L + thr * pm.emission;                       // :196
                    P.L[0][pid] = L.x; P.L[1][pid] = L.y; P.L[2][pid] = L.z;
                }
            }
            // None: `break` before the emission is added (src/renderer.rs:190-193)
        }
#if HNM_SHADE_OCTANT_SORT == 2
        uint32_t q2 = queue_alloc_warp_grouped(alive, direction_octant(nd), &P.counters[(bounce + 1) * C_STRIDE + C_RAY]);
#elif HNM_SHADE_OCTANT_SORT
        uint32_t q2 = queue_alloc_grouped(alive, direction_octant(nd), &P.counters[(bounce + 1) * C_STRIDE + C_RAY]);
#else
        uint32_t q2 = queue_alloc(alive, &P.counters[(bounce + 1) * C_STRIDE + C_RAY]);
#endif
        if (alive) store_ray(P, q2, no, nd, nthr, pid);
        if (NEE) {
            uint32_t ev = queue_alloc(event, &P.counters[bounce * C_STRIDE + C_EVENTS]);
            if (event) {
                P.ev_thr[0][ev] = thr.x; P.ev_thr[1][ev] = thr.y; P.ev_thr[2][ev] = thr.z;
                P.ev_albedo[0][ev] = pm.albedo.x; P.ev_albedo[1][ev] = pm.albedo.y; P.ev_albedo[2][ev] = pm.albedo.z;
                P.ev_emission[0][ev] = pm.emission.x; P.ev_emission[1][ev] = pm.emission.y; P.ev_emission[2][ev] = pm.emission.z;
                P.ev_pid[ev] = pid;
                shadow_rays += nl;
                // next_event_estimation (src/renderer.rs:275-281): position = result.ray.origin
                for (uint32_t k = 0; k < nl; k++) {
                    const DElement& e = P.sc.elements[P.sc.emissions[k]];
                    // Sphere::sample_on_surface; theta = PI2 * random.0 is the phi of the BSDF sample
                    double unit_z = 1.0 - 2.0 * random.r1;
                    double a = __dsqrt_rn(1.0 - unit_z * unit_z);
                    D3 s_normal = d3(a * cos_phi, a * sin_phi, unit_z);
                    D3 s_position = d3(e.ax, e.ay, e.az) + (e.radius + P.sc.offset) * s_normal;
                    D3 shadow_vec = s_position - no;
                    D3 shadow_dir = normalize(shadow_vec);
                    double dot_0 = fabs(dot(sp.normal, shadow_dir));
                    double dot_l = fabs(dot(s_normal, shadow_dir));
                    double distance_pow2 = dot(shadow_vec, shadow_vec);
                    size_t s = (size_t)ev * nl + k;
                    P.sray[0][s] = no.x; P.sray[1][s] = no.y; P.sray[2][s] = no.z;
                    P.sray[3][s] = shadow_dir.x; P.sray[4][s] = shadow_dir.y; P.sray[5][s] = shadow_dir.z;
                    P.s_pos[0][s] = s_position.x; P.s_pos[1][s] = s_position.y; P.s_pos[2][s] = s_position.z;
                    P.s_g[s] = (dot_0 * dot_l) / distance_pow2;
                    P.s_tmax[s] = (float)__dsqrt_rn(distance_pow2);
                    P.s_bsdf[s] = bsdf(pm, view, sp.normal, shadow_dir);
                }
            }
        }
    }
    if (NEE) {
        for (int o = 16; o > 0; o >>= 1) shadow_rays += __shfl_xor_sync(0xFFFFFFFFu, shadow_rays, o);
        if ((threadIdx.x & 31) == 0 && shadow_rays) {
            atomicAdd(&P.stats[S_SHADOW], (unsigned long long)shadow_rays);
            atomicAdd(&P.counters[bounce * C_STRIDE + C_SHADOW], shadow_rays);
        }
    }
}

// src/renderer.rs:282-295 + :183,196 for the NEE events of one bounce, after k_trace listed the candidates of their
// shadow rays: exact closest hit of every shadow ray (confirm_ray -- its only consumer is right here, so no hit record
// goes through memory), visibility test, light contribution, radiance update.
template <bool STATS, bool FAST>
__global__ void __launch_bounds__(256, HNM_NEER_MIN_BLOCKS) k_nee_resolve(RParams P, CandLists cand, int bounce) {
    const uint32_t n = P.counters[bounce * C_STRIDE + C_EVENTS];
    const uint32_t nl = P.sc.num_emissions;
    const DScene& sc = P.sc;
    uint32_t n_prims = 0;
    uint32_t* const work = &P.counters[bounce * C_STRIDE + C_W_NEER];
#if HNM_DYN_NEER
    for (;;) {
        const uint32_t base = fetch_warp(work);
        if (base >= n) break;
        const uint32_t ev = base + (threadIdx.x & 31);
        if (ev >= n) continue;
#else
    for (uint32_t ev = blockIdx.x * blockDim.x + threadIdx.x; ev < n; ev += gridDim.x * blockDim.x) {
#endif
        D3 accumulation = splat(0.0);
        for (uint32_t k = 0; k < nl; k++) {
            size_t s = (size_t)ev * nl + k;
            const uint32_t slot = P.cap + (uint32_t)s;  // the shadow rays' lists follow the camera rays'
            const uint32_t cn = __ldcs(cand.n + slot);
            if (STATS && cn == CAND_OVERFLOW) atomicAdd(&P.stats[S_OVERFLOW], 1ull);
            if (cn == 0u || cn == CAND_OCCLUDED) continue;  // nothing near the light sample, or something in front of it
            const float ub = __ldcs(cand.ub + slot);
            const uint32_t cid0 = __ldcs(cand.id + slot);
            const float lo0 = __ldcs(cand.lo + slot);
            D3 o = d3(P.sray[0][s], P.sray[1][s], P.sray[2][s]);
            D3 d = d3(P.sray[3][s], P.sray[4][s], P.sray[5][s]);
            const Hit h = confirm_ray<STATS>(sc, cand, slot, cn, ub, cid0, lo0, o, d, n_prims);
            if (h.kind == LEAF_NONE) continue;
            D3 s_position = d3(P.s_pos[0][s], P.s_pos[1][s], P.s_pos[2][s]);
            D3 hit_pos = o + d * h.t;
            if (norm(hit_pos - s_position) < sc.offset * 4.0) {  // Vector3::approximately (src/vector.rs:89-91)
                uint32_t el = h.kind == LEAF_TRI ? sc.tri_elem[h.id] : h.id;
                const DMaterial& hm = sc.materials[sc.elements[el].material];
                D3 emission;
                if (hm.emission.image >= 0) {
                    SurfacePoint sp = surface_point<FAST>(sc, h, o, d, true);
                    emission = texture_sample<FAST>(sc, hm.emission, sp.u, sp.v);
                } else {
                    emission = d3(hm.emission.r, hm.emission.g, hm.emission.b);
                }
                const DElement& e = sc.elements[sc.emissions[k]];
                double pdf = 1.0 / (4.0 * HNM_PI * e.radius * e.radius);
                accumulation = accumulation + emission * P.s_bsdf[s] * P.s_g[s] / pdf;
            }
        }
        D3 nee = accumulation * d3(P.ev_albedo[0][ev], P.ev_albedo[1][ev], P.ev_albedo[2][ev]);
        D3 thr = d3(P.ev_thr[0][ev], P.ev_thr[1][ev], P.ev_thr[2][ev]);
        D3 emission = d3(P.ev_emission[0][ev], P.ev_emission[1][ev], P.ev_emission[2][ev]);
        uint32_t pid = P.ev_pid[ev];
        D3 L = d3(P.L[0][pid], P.L[1][pid], P.L[2][pid]);
        L = L + thr * nee;       // src/renderer.rs:183
        L = L + thr * emission;  // :196
        P.L[0][pid] = L.x; P.L[1][pid] = L.y; P.L[2][pid] = L.z;
    }
    if (STATS) {
        for (int o = 16; o > 0; o >>= 1) n_prims += __shfl_xor_sync(0xFFFFFFFFu, n_prims, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(&P.stats[S_PRIMS], (unsigned long long)n_prims);
    }
}

// DebugRenderer::calc_pixel (src/renderer.rs:116-139)
__global__ void __launch_bounds__(256) k_debug_shade(RParams P) {
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < P.N; q += gridDim.x * blockDim.x) {
        D3 o = load_ray_o(P, q), d = load_ray_d(P, q);
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
    const int slot = isaac_slot();
    if (slot < 0) return;
    uint64_t* mem = smem_isaac + slot;
    for (uint32_t p = blockIdx.x * ISAAC_PATHS + slot; p < n; p += gridDim.x * ISAAC_PATHS) {
        uint64_t* o = out + (size_t)p * count;
        isaac64_seed<ISAAC_PATHS, RNG_TAIL>(mem, seeds[4 * p], seeds[4 * p + 1], seeds[4 * p + 2], seeds[4 * p + 3], [&](int i, uint64_t v) {
            int j = 255 - i;
            if (j < (int)count) o[j] = v;
        });
    }
}
__global__ void __launch_bounds__(ISAAC_TM_THREADS, 1) k_isaac_batch_tm(const uint64_t* seeds, uint32_t n, uint32_t count, uint64_t* out) {
    extern __shared__ uint64_t smem_isaac[];
    // static distribution: pair w of CTA b takes groups (4 b + w), (4 b + w) + 4 gridDim, ...
    uint32_t next = (blockIdx.x * 4u + ((threadIdx.x >> 5) & 3u)) * ISAAC_TM_LANES;
    const uint32_t stride = gridDim.x * 4u * ISAAC_TM_LANES;
    isaac64_tmem_pipeline<RNG_TAIL>(
        smem_isaac,
        [&]() -> uint32_t {
            const uint32_t g = next < n ? next : ISAAC_NONE;
            next += stride;
            return g;
        },
        [&](uint32_t p, uint64_t& s0, uint64_t& s1, uint64_t& s2, uint64_t& s3) {
            p = p < n ? p : n - 1;
            s0 = seeds[4 * p]; s1 = seeds[4 * p + 1]; s2 = seeds[4 * p + 2]; s3 = seeds[4 * p + 3];
        },
        [&](uint32_t p, int i, uint64_t v) {
            const int j = 255 - i;
            if (p < n && j < (int)count) out[(size_t)p * count + j] = v;
        },
        [&](uint32_t) {});
}
__global__ void k_isaac_full_batch(const uint64_t* seeds, uint32_t n, uint32_t count, uint64_t* out) {
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        IsaacFull rng;
        rng.seed(seeds[4 * p], seeds[4 * p + 1], seeds[4 * p + 2], seeds[4 * p + 3]);
        for (uint32_t j = 0; j < count; j++) out[(size_t)p * count + j] = rng.next_u64();
    }
}
// material resolve of traced hits (BvhScene::intersect after the traversal, src/scene.rs:389-398)
struct RayPtrs { const double* p[6]; };
__global__ void k_hits_to_abi(DScene sc, RayPtrs ray, const double* hit_t, const double* hit_u, const double* hit_v,
                              const uint2* hit_id, uint32_t n, hnm_hit* hits) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        D3 o = d3(ray.p[0][i], ray.p[1][i], ray.p[2][i]), d = d3(ray.p[3][i], ray.p[4][i], ray.p[5][i]);
        Hit h;
        h.t = hit_t[i]; h.u = hit_u[i]; h.v = hit_v[i]; h.kind = hit_id[i].x; h.id = hit_id[i].y;
        hnm_hit out;
        memset(&out, 0, sizeof(out));
        // Evaluated for every ray, not only for misses: with the call inside the `if`, ptxas 12.9 (sm_100a, -O3) returns
        // a wrong THIRD component of sample_bilinear's result to this kernel (PTX is correct; reproduced with an
        // all-miss input, fixed by hoisting the call).  This is a batch / test entry point, so the extra work is
        // irrelevant; the production kernels are pinned bit-for-bit by tests/test_gpu_parity.py.
        D3 e = skybox_sample(sc, d);
        if (h.kind == LEAF_NONE) {
            // Intersection::empty() + skybox emission (src/scene.rs:26-39,398)
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
__global__ void k_texture_sample_batch(DScene sc, DTexture t, const double* uv, uint32_t n, double* out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        D3 c = texture_sample(sc, t, uv[2 * (size_t)i], uv[2 * (size_t)i + 1]);
        out[3 * (size_t)i] = c.x; out[3 * (size_t)i + 1] = c.y; out[3 * (size_t)i + 2] = c.z;
    }
}
__global__ void k_skybox_sample_batch(DScene sc, const double* dirs, uint32_t n, double* out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        D3 c = skybox_sample(sc, d3(dirs[3 * (size_t)i], dirs[3 * (size_t)i + 1], dirs[3 * (size_t)i + 2]));
        out[3 * (size_t)i] = c.x; out[3 * (size_t)i + 1] = c.y; out[3 * (size_t)i + 2] = c.z;
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
#endif
