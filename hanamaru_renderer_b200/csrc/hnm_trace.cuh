// hnm_trace.cuh -- the closest-hit kernel: persistent warps, while-while traversal, dynamic ray fetch,
// f32 traversal with exact f64 confirmation.
//
// One launch serves up to two ray lists ("jobs"): the camera-path rays of bounce b and the NEE shadow rays of
// bounce b-1 (both are closest-hit queries in the reference: src/renderer.rs:176,280).  Each lane owns one ray.
//
// What decides a result is ONLY the reference's f64 primitive arithmetic (hnm_device.cuh: tri_test, sphere_test,
// cuboid_test); everything that merely decides WHICH primitives get that test runs in f32 and is conservative:
//   * node boxes: f32, rounded outward and padded (hnm_scene.cuh);
//   * triangles: a Moeller-Trumbore pre-test in f32 with running error bounds.  A triangle that may be hit at or
//     before the current bound becomes a CANDIDATE (kept in a 4-entry per-lane list); a triangle that is hit
//     for certain, by the full error margin, lowers `best_ub`, a safe upper bound of the closest-hit distance
//     that culls nodes and later candidates.  The exact f64 test runs on the surviving candidates when the ray
//     retires -- a warp-converged point, so the expensive f64 code executes with most lanes active -- or
//     earlier if a lane's list overflows.  ncu, round 1: 29.5 exact triangle tests per ray made FP64 the
//     busiest pipe at 5-18 of 32 lanes per instruction.
//   * spheres / cuboids (a handful per scene) are tested exactly on the spot.
// A warp keeps traversing until fewer than TRACE_REFILL lanes still have work; then finished lanes confirm
// their candidates, write their hits, push the hit's queue class (miss / delta BSDF / NEE BSDF; one atomic per
// warp and class) and pull new rays from a global work counter.
#ifndef HNM_TRACE_CUH
#define HNM_TRACE_CUH

#include "hnm_device.cuh"

namespace hnm {

#ifndef HNM_TRACE_MIN_BLOCKS
#define HNM_TRACE_MIN_BLOCKS 5  /* 96 registers: measured best of 4 (108 regs) / 5 / 6 (80 regs, spills) */
#endif
#ifndef HNM_TRACE_REFILL
#define HNM_TRACE_REFILL 20
#endif
#ifndef HNM_TRACE_TOPREG
#define HNM_TRACE_TOPREG 0      /* keep the stack top in a register */
#endif
#ifndef HNM_TRACE_NODE_STEPS
#define HNM_TRACE_NODE_STEPS 3  /* max node steps per scheduling vote */
#endif
#ifndef HNM_TRACE_LEAF_STEPS
#define HNM_TRACE_LEAF_STEPS 2  /* max leaf steps per scheduling vote */
#endif

constexpr int TRACE_THREADS = 128;
constexpr int TRACE_REFILL = HNM_TRACE_REFILL;  // refetch when fewer lanes than this are still traversing

struct TraceJob {
    const double* ray[6];  // origin xyz, direction xyz (SoA)
    double* hit_t; double* hit_u; double* hit_v; uint2* hit_id;
    const uint32_t* count;     // rays in this list (device memory)
    const float* tmax;         // optional: bounded query, see k_trace (null = closest hit along the whole ray)
    // classification of camera-path hits into shading queues (null for shadow rays)
    uint32_t* cnt_miss; uint32_t* cnt_delta; uint32_t* cnt_nee;
    uint32_t* q_miss; uint32_t* q_delta; uint32_t* q_nee;
};
struct TraceArgs {
    TraceJob job[2];
    uint32_t* work;                 // global fetch counter, zero at launch
    unsigned long long* stats;      // S_* counters
    int njobs;
    int stat_segments;              // stats index that receives job[0]'s ray count, or -1
    int stat_nodes, stat_prims;
    float tmax_slack;               // half-width of the window around tmax[] inside which the exact closest hit matters
};

HNM_D void warp_queue_push(bool pred, uint32_t* counter, uint32_t* queue, uint32_t value, int lane) {
    unsigned mask = __ballot_sync(0xFFFFFFFFu, pred);
    if (mask == 0) return;
    int leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(mask));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    if (pred) queue[base + __popc(mask & ((1u << lane) - 1u))] = value;
}

HNM_D float l1(float x, float y, float z) { return fabsf(x) + fabsf(y) + fabsf(z); }

// per-lane traversal state that the f32 phase reads
struct RayF {
    float ox, oy, oz;      // origin (advanced to the scene box if it was far away)
    float ix, iy, iz;      // 1 / direction
    float rx, ry, rz;      // -direction
    float K;               // absolute error bound of (origin - vertex) in f32
    float t0;              // distance the origin was advanced by (>= true value), 0 normally
};

// Conservative f32 triangle pre-test.  Returns false only if the exact test (src/bvh.rs:266-290) cannot accept
// the triangle at a distance <= best_ub.  May lower best_ub when the hit is certain.  *t_lo = lower bound of t.
HNM_D bool tri_pretest(const float4* __restrict__ tf, const RayF& R, float& best_ub, float* t_lo) {
    const float4 f0 = __ldg(tf), f1 = __ldg(tf + 1), f2 = __ldg(tf + 2);
    const float e1x = f0.w, e1y = f1.x, e1z = f1.y, e2x = f1.z, e2y = f1.w, e2z = f2.x;
    const float dx = R.ox - f0.x, dy = R.oy - f0.y, dz = R.oz - f0.z;
    // q = e2 x r,  den = e1 . q = det(e1, e2, r)
    const float qx = e2y * R.rz - e2z * R.ry, qy = e2z * R.rx - e2x * R.rz, qz = e2x * R.ry - e2y * R.rx;
    const float den = e1x * qx + e1y * qy + e1z * qz;
    const float sgn = den < 0.0f ? -1.0f : 1.0f;
    const float aden = fabsf(den);
    const float E1 = l1(e1x, e1y, e1z), E2 = l1(e2x, e2y, e2z), Q = l1(qx, qy, qz);
    const float KA = R.K + l1(dx, dy, dz) * 1.9073486328125e-06f;  // 2^-19: relative rounding of the products
    const float Mden = 1.9073486328125e-06f * E1 * (Q + E2);
    const float un = (dx * qx + dy * qy + dz * qz) * sgn;            // u * |den|
    const float Mu = KA * (Q + E2);
    if (un < -Mu || un > aden + Mu + Mden) return false;
    // p = r x e1,  v * den = d . p = det(e1, d, r)
    const float px = R.ry * e1z - R.rz * e1y, py = R.rz * e1x - R.rx * e1z, pz = R.rx * e1y - R.ry * e1x;
    const float vn = (dx * px + dy * py + dz * pz) * sgn;
    const float Mv = KA * (l1(px, py, pz) + E1);
    if (vn < -Mv || un + vn > aden + Mu + Mv + Mden) return false;
    // t * den = d . (e1 x e2)
    const float tn = (dx * f2.y + dy * f2.z + dz * f2.w) * sgn;
    const float Mt = KA * l1(f2.y, f2.z, f2.w);
    const float dhi = aden + Mden;
    if (tn + Mt < -R.t0 * dhi) return false;              // t < 0 for sure (t is measured from the advanced origin)
    const float num_lo = tn - Mt;
    if (num_lo > best_ub * dhi) return false;             // t > best_ub for sure
    *t_lo = num_lo > 0.0f ? (num_lo / dhi) * 0.99999976f : -3.0e38f;
    const float dlo = aden - Mden;
    if (dlo > Mden && un >= Mu && vn >= Mv && un + vn <= dlo - (Mu + Mv) && num_lo >= 0.0f && R.t0 == 0.0f) {
        // inside by the full margin: the exact test accepts it at t <= (tn + Mt) / (aden - Mden)
        float t_ub = ((tn + Mt) / dlo) * 1.0000005f;
        best_ub = fminf(best_ub, t_ub);
    }
    return true;
}

// `cur` of a lane without a ray: a LEAF_NONE link, so that "holds a node" is cur >= 0 and "holds a leaf" is
// cur < 0 && cur != TRACE_IDLE -- two ballots per scheduling round instead of three
constexpr int32_t TRACE_IDLE = ~(int32_t)((uint32_t)LEAF_NONE << 29);

template <bool STATS>
__global__ void __launch_bounds__(TRACE_THREADS, HNM_TRACE_MIN_BLOCKS) k_trace(DScene sc, TraceArgs A) {
    const int lane = threadIdx.x & 31;
    const uint32_t n0 = *A.job[0].count;
    const uint32_t n1 = A.njobs > 1 ? *A.job[1].count : 0u;
    const uint32_t ntot = n0 + n1;
    const float W = 4.76837158203125e-07f;  // 2^-21, see hnm_device.cuh: trace()

    // traversal stack: the top entry lives in a register (`top`), the rest in local memory -- a pop hands `top` to
    // `cur` at once and the reload of the next entry is off the critical path (ncu, round 1: the dependent local load
    // `cur = stack[--sp]` in front of the node fetch was the most-sampled line of the kernel)
    int32_t stack[HNM_STACK];
    int sp = 0;          // entries on the stack INCLUDING `top`
    int32_t top = 0;
    int32_t cur = TRACE_IDLE;
    bool pending = false;
    uint32_t idx = 0;
    D3 o = splat(0.0), dir = splat(0.0);
    RayF R;
    R.ox = R.oy = R.oz = R.ix = R.iy = R.iz = R.rx = R.ry = R.rz = 0.f; R.K = 0.f; R.t0 = 0.f;
    double t0 = 0.0;
    float best_ub = 3.0e38f;
    float t_occ = -3.0e38f;  // shadow rays: a certain hit closer than this ends the ray (see TraceJob::tmax)
    Hit best;
    best.t = sc.inf; best.u = 0.0; best.v = 0.0; best.kind = LEAF_NONE; best.id = 0;
    // candidate triangles awaiting the exact test (newest first)
    uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    float l0 = 0.f, l1_ = 0.f, l2 = 0.f, l3 = 0.f;
    int ncand = 0;
    uint32_t n_nodes = 0, n_prims = 0;
    bool more = true;  // warp-uniform: the work counter has not run past the end yet
    uint32_t lk = 0;   // next primitive of the leaf this lane holds

#define HNM_EXACT(g)                                               \
    {                                                              \
        DTri tr_ = load_tri(sc.tris + (g));                        \
        if (STATS) n_prims++;                                      \
        tri_test(tr_, (g), o, dir, best);                          \
    }
#if HNM_TRACE_TOPREG
#define HNM_POP()                                                  \
    {                                                              \
        if (sp == 0) { cur = TRACE_IDLE; pending = true; }         \
        else {                                                     \
            cur = top;                                             \
            sp--;                                                  \
            if (sp > 0) top = stack[sp - 1];                       \
        }                                                          \
    }
#define HNM_PUSH(v)                                                \
    {                                                              \
        if (sp > 0) stack[sp - 1] = top;                           \
        top = (v);                                                 \
        sp++;                                                      \
    }
#else
#define HNM_POP()                                                  \
    {                                                              \
        if (sp == 0) { cur = TRACE_IDLE; pending = true; }         \
        else cur = stack[--sp];                                    \
    }
#define HNM_PUSH(v) { stack[sp++] = (v); }
#endif

    for (;;) {
        // ---- converged: confirm candidates of finished rays in f64, retire them ----------------------------
        {
            int cls = -1;
            if (pending) {
                // the exact closest hit is among the candidates whose lower bound does not exceed the bound
                if (ncand > 0 && l0 <= best_ub) HNM_EXACT(c0)
                if (ncand > 1 && l1_ <= best_ub) HNM_EXACT(c1)
                if (ncand > 2 && l2 <= best_ub) HNM_EXACT(c2)
                if (ncand > 3 && l3 <= best_ub) HNM_EXACT(c3)
                if (ncand < 0) { best.t = sc.inf; best.u = 0.0; best.v = 0.0; best.kind = LEAF_NONE; best.id = 0; }  // occluded shadow ray
                ncand = 0;
                const bool j1 = idx >= n0;
                const TraceJob& J = A.job[j1 ? 1 : 0];
                const uint32_t q = j1 ? idx - n0 : idx;
                __stcs(J.hit_t + q, best.t); __stcs(J.hit_u + q, best.u); __stcs(J.hit_v + q, best.v);
                __stcs(J.hit_id + q, make_uint2(best.kind, best.id));
                if (!j1 && J.q_miss) {
                    if (best.kind == LEAF_NONE) cls = 0;
                    else {
                        uint32_t el = best.kind == LEAF_TRI ? sc.tri_elem[best.id] : best.id;
                        int surface = sc.materials[sc.elements[el].material].surface;
                        cls = nee_available(surface) ? 2 : 1;
                    }
                }
                pending = false;
            }
            const TraceJob& J0 = A.job[0];
            if (J0.q_miss) {
                warp_queue_push(cls == 0, J0.cnt_miss, J0.q_miss, idx, lane);
                warp_queue_push(cls == 1, J0.cnt_delta, J0.q_delta, idx, lane);
                warp_queue_push(cls == 2, J0.cnt_nee, J0.q_nee, idx, lane);
            }
        }
        // ---- converged: fetch new rays -----------------------------------------------------------------------
        {
            unsigned need = __ballot_sync(0xFFFFFFFFu, cur == TRACE_IDLE);
            if (need) {
                int leader = __ffs(need) - 1;
                uint32_t base = 0;
                if (lane == leader) base = atomicAdd(A.work, (uint32_t)__popc(need));
                base = __shfl_sync(0xFFFFFFFFu, base, leader);
                more = base + (uint32_t)__popc(need) < ntot;
                if (cur == TRACE_IDLE) {
                    idx = base + __popc(need & ((1u << lane) - 1u));
                    if (idx < ntot) {
                        const bool j1 = idx >= n0;
                        const TraceJob& J = A.job[j1 ? 1 : 0];
                        const uint32_t q = j1 ? idx - n0 : idx;
                        // streaming (evict-first) loads and stores for the ray / hit records: the L1 is for the tree
                        o = d3(__ldcs(J.ray[0] + q), __ldcs(J.ray[1] + q), __ldcs(J.ray[2] + q));
                        dir = d3(__ldcs(J.ray[3] + q), __ldcs(J.ray[4] + q), __ldcs(J.ray[5] + q));
                        best.t = sc.inf; best.u = 0.0; best.v = 0.0; best.kind = LEAF_NONE; best.id = 0;
                        best_ub = 3.0e38f;
                        t_occ = -3.0e38f;
                        ncand = 0;
                        t0 = 0.0;
                        sp = 0;
                        cur = 0;
                        lk = 0;
                        float fmaxo = fmaxf(fmaxf(fabsf((float)o.x), fabsf((float)o.y)), fabsf((float)o.z));
                        if (fmaxo > sc.far_limit) {
                            double dist;
                            bool h = aabb_intersect_ray(sc.bounds_lo[0], sc.bounds_lo[1], sc.bounds_lo[2], sc.bounds_hi[0], sc.bounds_hi[1],
                                                        sc.bounds_hi[2], o, dir, &dist);
                            if (!h) { cur = TRACE_IDLE; pending = true; }
                            else if (dist > 0.0 && dist < sc.inf) t0 = dist * (1.0 - 1e-6);
                        } else if (J.tmax) {
                            // Bounded query (NEE shadow rays): the caller only looks at the closest hit if it lies within
                            // `slack` of distance tmax[q] along the ray (src/renderer.rs:282, Vector3::approximately).
                            // Nothing beyond tmax + slack can be that hit, and a certain hit before tmax - slack means the
                            // closest hit is not it either: the ray ends at once and reports "no hit".
                            const float D = __ldcs(J.tmax + q);
                            best_ub = D + A.tmax_slack;
                            t_occ = D - A.tmax_slack;
                        }
                        R.ox = (float)(o.x + dir.x * t0); R.oy = (float)(o.y + dir.y * t0); R.oz = (float)(o.z + dir.z * t0);
                        R.ix = (float)(1.0 / dir.x); R.iy = (float)(1.0 / dir.y); R.iz = (float)(1.0 / dir.z);
                        R.rx = (float)(-dir.x); R.ry = (float)(-dir.y); R.rz = (float)(-dir.z);
                        R.t0 = __double2float_ru(t0);
                        // |f32(origin) - origin| + |f32(vertex) - vertex| + rounding of the difference, with 4x slack
                        R.K = (fmaxf(fmaxf(fabsf(R.ox), fabsf(R.oy)), fabsf(R.oz)) + sc.scene_r) * 4.76837158203125e-07f;
                    }
                }
            }
            if (__ballot_sync(0xFFFFFFFFu, cur != TRACE_IDLE || pending) == 0) break;
        }
        // ---- traverse (f32) until too few lanes are left -----------------------------------------------------
        // Majority-vote scheduling.  Each iteration every lane is in one of two states: it holds a NODE (cur >= 0)
        // or a LEAF with triangles left (cur < 0).  The warp executes ONE step of whichever kind the majority of
        // lanes needs -- a node step (two f32 slab tests) or a leaf step (one f32 triangle pre-test) -- and the
        // minority idles until it becomes the majority.  The loop is warp-uniform (full-mask ballots): with
        // independent thread scheduling a per-lane `while (has_ray)` never reconverges, and a strict while-while
        // loop (all lanes descend to a leaf, then all test it) ran at 4 of 32 lanes per instruction on the
        // incoherent bounces because the longest descent of the warp sets the pace (ncu, round 1).
        for (;;) {
            const bool is_node = cur >= 0;
            const unsigned nm = __ballot_sync(0xFFFFFFFFu, is_node);
            const unsigned lm = __ballot_sync(0xFFFFFFFFu, cur < 0 && cur != TRACE_IDLE);
            const int nn = __popc(nm), nl = __popc(lm);
            if (nn + nl == 0 || (more && nn + nl < TRACE_REFILL)) break;  // too few lanes left: refill (lanes keep their state)
            if (nn >= nl) {
                // node phase: HNM_TRACE_NODE_STEPS steps per vote, unrolled, no ballot in between (lanes that reach a
                // leaf early idle for the rest of the phase; re-voting after every step cost more than it saved)
#pragma unroll
                for (int rep = 0; rep < HNM_TRACE_NODE_STEPS; rep++) {
                    if (cur >= 0) {
                        const float4* np = reinterpret_cast<const float4*>(sc.nodes + cur);
                        float4 m0 = __ldg(np), m1 = __ldg(np + 1), m2 = __ldg(np + 2);
                        int2 m3 = __ldg(reinterpret_cast<const int2*>(np + 3));
                        if (STATS) n_nodes++;
                        float a0 = (m0.x - R.ox) * R.ix, b0 = (m0.w - R.ox) * R.ix;
                        float a1 = (m0.y - R.oy) * R.iy, b1 = (m1.x - R.oy) * R.iy;
                        float a2 = (m0.z - R.oz) * R.iz, b2 = (m1.y - R.oz) * R.iz;
                        float tmin0 = fmaxf(fmaxf(fminf(a0, b0), fminf(a1, b1)), fminf(a2, b2));
                        float tmax0 = fminf(fminf(fmaxf(a0, b0), fmaxf(a1, b1)), fmaxf(a2, b2));
                        float g0 = (m1.z - R.ox) * R.ix, e0 = (m2.y - R.ox) * R.ix;
                        float g1 = (m1.w - R.oy) * R.iy, e1 = (m2.z - R.oy) * R.iy;
                        float g2 = (m2.x - R.oz) * R.iz, e2 = (m2.w - R.oz) * R.iz;
                        float tmin1 = fmaxf(fmaxf(fminf(g0, e0), fminf(g1, e1)), fminf(g2, e2));
                        float tmax1 = fminf(fminf(fmaxf(g0, e0), fmaxf(g1, e1)), fmaxf(g2, e2));
                        float lo0 = tmin0 - fabsf(tmin0) * W, up0 = tmax0 + fabsf(tmax0) * W;
                        float lo1 = tmin1 - fabsf(tmin1) * W, up1 = tmax1 + fabsf(tmax1) * W;
                        const bool h0 = (lo0 <= up0) && (up0 >= 0.0f) && (lo0 <= best_ub);
                        const bool h1 = (lo1 <= up1) && (up1 >= 0.0f) && (lo1 <= best_ub);
                        const bool swap = lo1 < lo0;
                        lk = 0;
                        if (h0 || h1) {
                            // near child next; the far one (if both are hit) becomes the new stack top
                            cur = (h0 && !(h1 && swap)) ? m3.x : m3.y;
                            if (h0 && h1 && sp < HNM_STACK) HNM_PUSH(swap ? m3.x : m3.y)
                        } else {
                            HNM_POP()
                        }
                    }
                }
            } else {
#pragma unroll
              for (int rep = 0; rep < HNM_TRACE_LEAF_STEPS; rep++) {
                if (cur < 0 && cur != TRACE_IDLE) {
                // one primitive of the leaf this lane holds
                const int kind = leaf_kind(cur);
                const uint32_t first = leaf_first(cur);
                bool leaf_done = true;
                if (kind == LEAF_TRI) {
                    float tlo;
                    const uint32_t pos = first + lk;
                    if (tri_pretest(sc.trif + 3 * (size_t)pos, R, best_ub, &tlo)) {
                        const uint32_t g = __ldg(sc.tri_perm + pos);  // index in the reference's leaf order (ties)
                        if (ncand == 4) {
                            // list full: the oldest entry either drops out (bound moved below it) or is confirmed now
                            if (l3 <= best_ub) {
                                HNM_EXACT(c3)
                                best_ub = fminf(best_ub, __double2float_ru(best.t - t0));
                            }
                            ncand = 3;
                        }
                        c3 = c2; l3 = l2; c2 = c1; l2 = l1_; c1 = c0; l1_ = l0;
                        c0 = g; l0 = tlo;
                        ncand++;
                    }
                    lk++;
                    leaf_done = lk >= leaf_count(cur);
                } else if (kind == LEAF_SPHERE) {
                    if (STATS) n_prims++;
                    sphere_test(sc.elements[first], first, sc.elements, o, dir, best);
                    best_ub = fminf(best_ub, __double2float_ru(best.t - t0));
                } else if (kind == LEAF_CUBOID) {
                    if (STATS) n_prims++;
                    cuboid_test(sc.elements[first], first, sc.elements, o, dir, best);
                    best_ub = fminf(best_ub, __double2float_ru(best.t - t0));
                }
                if (best_ub < t_occ) {
                    // bounded query: something certainly lies in front of the point the caller asked about
                    ncand = -1;
                    cur = TRACE_IDLE;
                    pending = true;
                } else if (leaf_done) {
                    lk = 0;
                    HNM_POP()
                }
                }
              }
            }
        }
    }
#undef HNM_EXACT
#undef HNM_POP
#undef HNM_PUSH
    if (blockIdx.x == 0 && threadIdx.x == 0 && A.stat_segments >= 0) atomicAdd(&A.stats[A.stat_segments], (unsigned long long)n0);
    if (STATS) {
        for (int s = 16; s > 0; s >>= 1) { n_nodes += __shfl_xor_sync(0xFFFFFFFFu, n_nodes, s); n_prims += __shfl_xor_sync(0xFFFFFFFFu, n_prims, s); }
        if (lane == 0) { atomicAdd(&A.stats[A.stat_nodes], (unsigned long long)n_nodes); atomicAdd(&A.stats[A.stat_prims], (unsigned long long)n_prims); }
    }
}

}  // namespace hnm
#endif
