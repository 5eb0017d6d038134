// hnm_trace.cuh -- closest hit in two kernels: k_trace (f32 candidate search) and k_confirm (exact f64 tests).
//
// What decides a result is ONLY the reference's f64 primitive arithmetic (hnm_device.cuh: tri_test, sphere_test,
// cuboid_test).  Everything that merely decides WHICH primitives get that test is conservative f32 work:
//   * node boxes: f32, rounded outward and padded (hnm_scene.cuh);
//   * triangles: a Moeller-Trumbore pre-test in f32 with running error bounds; spheres and cuboids: f32 versions of
//     their own tests with error bounds.  A primitive that may be hit at or before the current bound becomes a
//     CANDIDATE of its ray (a list of up to TRACE_CAND entries in global memory); a primitive that is hit for
//     certain, by the full error margin, lowers `best_ub`, a safe upper bound of the closest-hit distance that
//     culls nodes and later candidates.
// k_trace: persistent warps, dynamic ray fetch, majority-vote scheduled traversal.  It holds no f64 state at all
// (ncu, round 1: with the exact tests and the f64 ray inside this kernel it needed 96-110 registers, 5 CTAs/SM,
// and 4 -> 5 CTAs/SM alone was worth 17 %).  It writes per ray: the candidate list, its length (or a flag) and
// the final bound.
// k_confirm: one thread per ray, coalesced: the exact f64 test on the candidates whose lower bound does not
// exceed the final bound -> the closest hit (ties resolved by the reference's DFS order, see tri_test), written
// as the hit record; camera-path hits are classified into the miss / delta-BSDF / NEE-BSDF shading queues.  A ray
// whose list overflowed is traced again here by the plain exact traversal (hnm_device.cuh: trace()).
// One launch of each serves up to two ray lists ("jobs"): the camera-path rays of bounce b and the NEE shadow rays
// of bounce b-1 (both are closest-hit queries in the reference: src/renderer.rs:176,280).
#ifndef HNM_TRACE_CUH
#define HNM_TRACE_CUH

#include "hnm_device.cuh"

namespace hnm {

#ifndef HNM_TRACE_MIN_BLOCKS
#define HNM_TRACE_MIN_BLOCKS 8  /* 64 registers */
#endif
#ifndef HNM_TRACE_REFILL
#define HNM_TRACE_REFILL 20
#endif
#ifndef HNM_TRACE_NODE_STEPS
#define HNM_TRACE_NODE_STEPS 3  /* max node steps per scheduling vote */
#endif
#ifndef HNM_TRACE_LEAF_STEPS
#define HNM_TRACE_LEAF_STEPS 2  /* max leaf steps per scheduling vote */
#endif

#ifndef HNM_TRACE_FMA_SLABS
#define HNM_TRACE_FMA_SLABS 1
#endif
#ifndef HNM_CONFIRM_MIN_BLOCKS
#define HNM_CONFIRM_MIN_BLOCKS 3
#endif
constexpr int TRACE_THREADS = 128;
constexpr int TRACE_REFILL = HNM_TRACE_REFILL;  // refetch when fewer lanes than this are still traversing

#ifndef HNM_TRACE_CAND
#define HNM_TRACE_CAND 32
#endif
constexpr int TRACE_CAND = HNM_TRACE_CAND;       // candidate-list capacity per ray (<= 32: k_confirm_pairs keeps a bit per entry)
static_assert(TRACE_CAND >= 1 && TRACE_CAND <= 32, "candidate-list capacity");
constexpr uint32_t CAND_OVERFLOW = 0xFFFFFFFFu;  // list overflowed: k_confirm runs the exact traversal
constexpr uint32_t CAND_OCCLUDED = 0xFFFFFFFEu;  // bounded query ended early: report "no hit"
// candidate id: kind (LEAF_TRI / LEAF_SPHERE / LEAF_CUBOID) << 30 | triangle index (reference leaf order) or element id

struct TraceJob {
    const double* ray[6];  // origin xyz, direction xyz (SoA)
    const uint32_t* count;     // rays in this list (device memory)
    const float* tmax;         // optional: bounded query, see k_trace (null = closest hit along the whole ray)
    uint32_t slot0;            // candidate-list slot of this job's ray 0
    // written by k_confirm
    double* hit_t; double* hit_u; double* hit_v; uint2* hit_id;
    // classification of camera-path hits into shading queues (null for shadow rays)
    uint32_t* cnt_miss; uint32_t* cnt_delta; uint32_t* cnt_nee;
    uint32_t* q_miss; uint32_t* q_delta; uint32_t* q_nee;
};
struct CandLists {
    uint32_t* id;   // [TRACE_CAND][stride]
    float* lo;      // [TRACE_CAND][stride] lower bound of the candidate's distance
    uint32_t* n;    // [stride] entries, or CAND_OVERFLOW / CAND_OCCLUDED
    float* ub;      // [stride] final upper bound of the closest-hit distance
    uint32_t stride;
};
struct TraceArgs {
    TraceJob job[2];
    CandLists cand;
    uint32_t* work;                 // global fetch counter, zero at launch
    uint32_t* work_confirm;         // the same for k_confirm
    unsigned long long* stats;      // S_* counters
    int njobs;
    int stat_segments;              // stats index that receives job[0]'s ray count, or -1
    int stat_nodes, stat_prims;
    float tmax_slack;               // half-width of the window around tmax[] inside which the exact closest hit matters
    unsigned long long* dbg;        // diagnostics: see RParams::dbg
};

HNM_D float l1(float x, float y, float z) { return fabsf(x) + fabsf(y) + fabsf(z); }

// per-lane traversal state that the f32 phase reads
struct RayF {
    float ox, oy, oz;      // origin (advanced to the scene box if it was far away)
    float ix, iy, iz;      // 1 / direction
    float rx, ry, rz;      // -direction
    float K;               // absolute error bound of (origin - vertex) in f32
    float t0;              // distance the origin was advanced by (>= true value), 0 normally
#if HNM_TRACE_FMA_SLABS
    // node slab tests as one fma per plane: t = plane * inv + c with c = -(origin * inv), widened by the rounding of
    // that product (2^-22 relative to |origin * inv|) towards -inf for the entry planes and +inf for the exit planes
    float nx, ny, nz, fx, fy, fz;
#endif
};

// Conservative f32 triangle pre-test.  Returns false only if the exact test (src/bvh.rs:266-290) cannot accept
// the triangle at a distance <= best_ub.  May lower best_ub when the hit is certain.  *t_lo = lower bound of t.
HNM_D bool tri_pretest(const float4* __restrict__ tf, const RayF& R, float& best_ub, float* t_lo) {
    const float4 f0 = __ldg(tf), f1 = __ldg(tf + 1), f2 = __ldg(tf + 2);
    const float e1x = f0.w, e1y = f1.x, e1z = f1.y, e2x = f1.z, e2y = f1.w, e2z = f2.x;
    const float dx = R.ox - f0.x, dy = R.oy - f0.y, dz = R.oz - f0.z;
    // q = e2 x r,  den = e1 . q = det(e1, e2, r)
    // (explicit fma: fewer instructions and fewer roundings than the error bounds below allow for)
    const float qx = __fmaf_rn(e2y, R.rz, -(e2z * R.ry)), qy = __fmaf_rn(e2z, R.rx, -(e2x * R.rz)), qz = __fmaf_rn(e2x, R.ry, -(e2y * R.rx));
    const float den = __fmaf_rn(e1x, qx, __fmaf_rn(e1y, qy, e1z * qz));
    const float sgn = den < 0.0f ? -1.0f : 1.0f;
    const float aden = fabsf(den);
    const float E1 = l1(e1x, e1y, e1z), E2 = l1(e2x, e2y, e2z), Q = l1(qx, qy, qz);
    const float KA = R.K + l1(dx, dy, dz) * 1.9073486328125e-06f;  // 2^-19: relative rounding of the products
    const float Mden = 1.9073486328125e-06f * E1 * (Q + E2);
    const float un = __fmaf_rn(dx, qx, __fmaf_rn(dy, qy, dz * qz)) * sgn;  // u * |den|
    const float Mu = KA * (Q + E2);
    if (un < -Mu || un > aden + Mu + Mden) return false;
    // p = r x e1,  v * den = d . p = det(e1, d, r)
    const float px = __fmaf_rn(R.ry, e1z, -(R.rz * e1y)), py = __fmaf_rn(R.rz, e1x, -(R.rx * e1z)), pz = __fmaf_rn(R.rx, e1y, -(R.ry * e1x));
    const float vn = __fmaf_rn(dx, px, __fmaf_rn(dy, py, dz * pz)) * sgn;
    const float Mv = KA * (l1(px, py, pz) + E1);
    if (vn < -Mv || un + vn > aden + Mu + Mv + Mden) return false;
    // t * den = d . (e1 x e2)
    const float tn = __fmaf_rn(dx, f2.y, __fmaf_rn(dy, f2.z, dz * f2.w)) * sgn;
    const float Mt = KA * l1(f2.y, f2.z, f2.w);
    const float dhi = aden + Mden;
    if (tn + Mt < -R.t0 * dhi) return false;              // t < 0 for sure (t is measured from the advanced origin)
    const float num_lo = tn - Mt;
    if (num_lo > best_ub * dhi) return false;             // t > best_ub for sure
    // (approximate division, 2 ulp, widened by 2^-20: the quotient only has to be a bound)
    *t_lo = num_lo > 0.0f ? __fdividef(num_lo, dhi) * 0.99999905f : -3.0e38f;
    const float dlo = aden - Mden;
    if (dlo > Mden && un >= Mu && vn >= Mv && un + vn <= dlo - (Mu + Mv) && num_lo >= 0.0f && R.t0 == 0.0f) {
        // inside by the full margin: the exact test accepts it at t <= (tn + Mt) / (aden - Mden)
        float t_ub = __fdividef(tn + Mt, dlo) * 1.000001f;
        best_ub = fminf(best_ub, t_ub);
    }
    return true;
}

// Conservative f32 version of Cuboid::intersect's slab test (src/scene.rs:152-183 via src/bvh.rs:20-39).
// ef[0], ef[1] = the box rounded OUTWARD and padded, ef[2], ef[3] = rounded INWARD and shrunk by the same pad.
// Returns false only if the exact test cannot hit at a distance <= best_ub; *t_lo = lower bound of that distance.
// Lowers best_ub when the ray passes through the inner box for certain.
HNM_D bool cuboid_pretest(const float4* __restrict__ ef, const RayF& R, float& best_ub, float* t_lo) {
    const float W = 4.76837158203125e-07f;  // 2^-21
    const float4 lo = __ldg(ef), hi = __ldg(ef + 1);
    float a0 = (lo.x - R.ox) * R.ix, b0 = (hi.x - R.ox) * R.ix;
    float a1 = (lo.y - R.oy) * R.iy, b1 = (hi.y - R.oy) * R.iy;
    float a2 = (lo.z - R.oz) * R.iz, b2 = (hi.z - R.oz) * R.iz;
    float tmin = fmaxf(fmaxf(fminf(a0, b0), fminf(a1, b1)), fminf(a2, b2));
    float tmax = fminf(fminf(fmaxf(a0, b0), fmaxf(a1, b1)), fmaxf(a2, b2));
    const float out_lo = tmin - fabsf(tmin) * W, out_up = tmax + fabsf(tmax) * W;
    // distances are measured from the (possibly advanced) f32 origin: the true t is larger by t0 >= 0
    if (!((out_lo <= out_up) && (out_up >= -R.t0) && (out_lo <= best_ub))) return false;
    *t_lo = out_lo;
    if (R.t0 == 0.0f) {
        const float4 li = __ldg(ef + 2), hj = __ldg(ef + 3);
        a0 = (li.x - R.ox) * R.ix; b0 = (hj.x - R.ox) * R.ix;
        a1 = (li.y - R.oy) * R.iy; b1 = (hj.y - R.oy) * R.iy;
        a2 = (li.z - R.oz) * R.iz; b2 = (hj.z - R.oz) * R.iz;
        tmin = fmaxf(fmaxf(fminf(a0, b0), fminf(a1, b1)), fminf(a2, b2));
        tmax = fminf(fminf(fmaxf(a0, b0), fmaxf(a1, b1)), fmaxf(a2, b2));
        const float in_lo = tmin + fabsf(tmin) * W, in_up = tmax - fabsf(tmax) * W;
        // (li <= hj componentwise is the host's job: a degenerate inner box is stored as NaN and never passes)
        if (in_lo <= in_up && in_up > 0.0f) {
            // the ray crosses the inner box, so the exact test hits.  Its distance is tmin if tmin >= 0 (entry into the
            // true box, not later than the entry into the inner box), else tmax (<= out_up).
            const float t_ub = out_lo > 0.0f ? in_lo : out_up;
            best_ub = fminf(best_ub, t_ub);
        }
    }
    return true;
}

// Conservative f32 version of Sphere::intersect (src/scene.rs:58-78): t = -b - sqrt(b^2 - c), hit iff d > 0 && t > 0.
// ef[0] = (centre, radius).  Same contract as cuboid_pretest.
HNM_D bool sphere_pretest(const float4* __restrict__ ef, const RayF& R, float& best_ub, float* t_lo) {
    const float E = 9.5367431640625e-07f;  // 2^-20: a generous bound for a handful of f32 roundings
    const float4 c = __ldg(ef);
    const float ax = R.ox - c.x, ay = R.oy - c.y, az = R.oz - c.z;  // each within R.K of the exact difference
    const float dx = -R.rx, dy = -R.ry, dz = -R.rz;
    const float A1 = l1(ax, ay, az);
    const float b = ax * dx + ay * dy + az * dz;
    const float Eb = 2.0f * R.K + A1 * E;                           // |d|_1 <= sqrt(3) < 2
    const float aa = ax * ax + ay * ay + az * az;
    const float rr = c.w * c.w;
    const float cc = aa - rr;
    const float Ec = 2.0f * A1 * R.K + 3.0f * R.K * R.K + (aa + rr) * E;
    const float disc = b * b - cc;
    const float Ed = 2.0f * fabsf(b) * Eb + Eb * Eb + Ec + (b * b + fabsf(cc)) * E;
    if (!(disc + Ed > 0.0f)) return false;                          // d <= 0 for certain
    const float s_hi = sqrtf(disc + Ed) * 1.000001f;
    const float s_lo = disc - Ed > 0.0f ? sqrtf(disc - Ed) * 0.999999f : 0.0f;
    const float Er = (fabsf(b) + s_hi) * E;                         // roundings of the two subtractions below
    const float tl = ((-b - Eb) - s_hi) - Er, th = ((-b + Eb) - s_lo) + Er;  // tl <= t <= th (from the f32 origin)
    if (!(th > -R.t0)) return false;                                // t <= 0 for certain
    if (tl > best_ub) return false;
    *t_lo = tl;
    if (disc - Ed > 0.0f && tl > 0.0f && R.t0 == 0.0f) best_ub = fminf(best_ub, th * 1.000001f + 1e-30f);
    return true;
}

// The exact fallback for one ray: hnm_device.cuh's trace() with two additions.  (1) Every leaf triangle first takes the f32
// pre-test above and only the survivors the f64 test: a ray that overflows its candidate list grazes dense geometry, and
// trace() ran ~190 f64 triangle tests for it (ncu, 75 k-triangle scene: the single lane doing that set the duration of every
// k_confirm launch, ~1 ms each whatever the ray count).  (2) `ub_cull` -- the bound k_trace had reached when it gave up --
// culls nodes and triangles from the start (the listed path culls with the same bound: `lo > ub`).  The pre-test's own
// "certain hit" tightening is NOT used (a copy of the bound is passed): rays with infinite reciprocals come here precisely
// because a geometrically certain hit may be one the reference's box chain never reaches; only confirmed hits tighten.
template <bool STATS>
HNM_D Hit trace_pretested(const DScene& sc, D3 o, D3 dir, float ub_cull, TraceStats* st) {
    Hit best;
    best.t = sc.inf; best.u = 0.0; best.v = 0.0; best.kind = LEAF_NONE; best.id = 0;
    double t0 = 0.0;
    float fmaxo = fmaxf(fmaxf(fabsf((float)o.x), fabsf((float)o.y)), fabsf((float)o.z));
    if (fmaxo > sc.far_limit) {
        double dist;
        bool h = aabb_intersect_ray(sc.bounds_lo[0], sc.bounds_lo[1], sc.bounds_lo[2], sc.bounds_hi[0], sc.bounds_hi[1], sc.bounds_hi[2], o, dir, &dist);
        if (!h) return best;
        if (dist > 0.0 && dist < sc.inf) t0 = dist * (1.0 - 1e-6);
    }
    RayF R;
    R.ox = (float)(o.x + dir.x * t0); R.oy = (float)(o.y + dir.y * t0); R.oz = (float)(o.z + dir.z * t0);
    R.ix = (float)(1.0 / dir.x); R.iy = (float)(1.0 / dir.y); R.iz = (float)(1.0 / dir.z);
    R.rx = (float)(-dir.x); R.ry = (float)(-dir.y); R.rz = (float)(-dir.z);
    R.t0 = __double2float_ru(t0);
    R.K = (fmaxf(fmaxf(fabsf(R.ox), fabsf(R.oy)), fabsf(R.oz)) + sc.scene_r) * 4.76837158203125e-07f;  // as in k_trace
#if HNM_TRACE_FMA_SLABS
    R.nx = R.ny = R.nz = R.fx = R.fy = R.fz = 0.f;  // (the slab tests below use the subtract form)
#endif
    const float W = 4.76837158203125e-07f;  // 2^-21: covers the rounding of (lo-o)*inv in f32
    int32_t stack[HNM_STACK];
    int sp = 0;
    int32_t cur = 0;
    for (;;) {
        // distances are measured from the advanced origin; a confirmed hit or k_trace's bound culls
        const float bestf = fminf(__double2float_ru(best.t - t0), ub_cull);
        if (cur >= 0) {
            const float4* np = reinterpret_cast<const float4*>(sc.nodes + cur);
            float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2);
            int4 n3 = __ldg(reinterpret_cast<const int4*>(np + 3));
            if (STATS) st->nodes++;
            float a0 = (n0.x - R.ox) * R.ix, b0 = (n0.w - R.ox) * R.ix;
            float a1 = (n0.y - R.oy) * R.iy, b1 = (n1.x - R.oy) * R.iy;
            float a2 = (n0.z - R.oz) * R.iz, b2 = (n1.y - R.oz) * R.iz;
            float tmin0 = fmaxf(fmaxf(fminf(a0, b0), fminf(a1, b1)), fminf(a2, b2));
            float tmax0 = fminf(fminf(fmaxf(a0, b0), fmaxf(a1, b1)), fmaxf(a2, b2));
            float c0 = (n1.z - R.ox) * R.ix, e0 = (n2.y - R.ox) * R.ix;
            float c1 = (n1.w - R.oy) * R.iy, e1 = (n2.z - R.oy) * R.iy;
            float c2 = (n2.x - R.oz) * R.iz, e2 = (n2.w - R.oz) * R.iz;
            float tmin1 = fmaxf(fmaxf(fminf(c0, e0), fminf(c1, e1)), fminf(c2, e2));
            float tmax1 = fminf(fminf(fmaxf(c0, e0), fmaxf(c1, e1)), fmaxf(c2, e2));
            float lo0 = tmin0 - fabsf(tmin0) * W, up0 = tmax0 + fabsf(tmax0) * W;
            float lo1 = tmin1 - fabsf(tmin1) * W, up1 = tmax1 + fabsf(tmax1) * W;
            bool h0 = (lo0 <= up0) && (up0 >= 0.0f) && (lo0 <= bestf);
            bool h1 = (lo1 <= up1) && (up1 >= 0.0f) && (lo1 <= bestf);
            if (h0 && h1) {
                bool swap = lo1 < lo0;
                int32_t nearc = swap ? n3.y : n3.x;
                int32_t farc = swap ? n3.x : n3.y;
                if (sp < HNM_STACK) stack[sp++] = farc;
                cur = nearc;
                continue;
            } else if (h0) {
                cur = n3.x;
                continue;
            } else if (h1) {
                cur = n3.y;
                continue;
            }
        } else {
            int kind = leaf_kind(cur);
            uint32_t first = leaf_first(cur);
            if (kind == LEAF_TRI) {
                uint32_t cnt = leaf_count(cur);
                for (uint32_t k = 0; k < cnt; k++) {
                    float ubc = fminf(__double2float_ru(best.t - t0), ub_cull), tlo;
                    if (!tri_pretest(sc.trif + 3 * (size_t)(first + k), R, ubc, &tlo)) continue;  // the exact test cannot accept it
                    uint32_t g = __ldg(sc.tri_perm + first + k);
                    DTri tr = load_tri(sc.tris + g);
                    if (STATS) st->prims++;
                    tri_test(sc, tr, g, o, dir, best);
                }
            } else if (kind == LEAF_SPHERE) {
                if (STATS) st->prims++;
                sphere_test(sc, sc.elements[first], first, sc.elements, o, dir, best);
            } else if (kind == LEAF_CUBOID) {
                if (STATS) st->prims++;
                cuboid_test(sc, sc.elements[first], first, sc.elements, o, dir, best);
            }
        }
        if (sp == 0) break;
        cur = stack[--sp];
    }
    return best;
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
    const float WLO = 1.0f - 4.76837158203125e-07f, WUP = 1.0f + 4.76837158203125e-07f;  // 1 -+ 2^-21, see hnm_device.cuh: trace()
    if (A.dbg && lane == 0) {
        unsigned w;
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(w));
        atomicOr(&A.dbg[1], 1ull << (w & 63u));
    }

    int32_t stack[HNM_STACK];
    int sp = 0;
    int32_t cur = TRACE_IDLE;
    uint32_t slot = 0;       // this ray's candidate list
    RayF R;
    R.ox = R.oy = R.oz = R.ix = R.iy = R.iz = R.rx = R.ry = R.rz = 0.f; R.K = 0.f; R.t0 = 0.f;
#if HNM_TRACE_FMA_SLABS
    R.nx = R.ny = R.nz = R.fx = R.fy = R.fz = 0.f;
#endif
    float best_ub = 3.0e38f;
    float t_occ = -3.0e38f;  // shadow rays: a certain hit closer than this ends the ray (see TraceJob::tmax)
    uint32_t ncand = 0;
    uint32_t n_nodes = 0;
    bool more = true;  // warp-uniform: the work counter has not run past the end yet
    uint32_t lk = 0;   // next primitive of the leaf this lane holds

    // the ray is finished: publish the list header (divergent, but two 4-byte stores)
#define HNM_FINISH(count_or_flag)                                  \
    {                                                              \
        __stcs(A.cand.n + slot, (uint32_t)(count_or_flag));        \
        __stcs(A.cand.ub + slot, best_ub);                         \
        cur = TRACE_IDLE;                                          \
    }
#define HNM_POP()                                                  \
    {                                                              \
        if (sp == 0) HNM_FINISH(ncand)                             \
        else cur = stack[--sp];                                    \
    }

    for (;;) {
        // ---- converged: fetch new rays -----------------------------------------------------------------------
        {
            unsigned need = __ballot_sync(0xFFFFFFFFu, cur == TRACE_IDLE);
            if (need) {
                int leader = __ffs(need) - 1;
                uint32_t base = 0;
                if (more && lane == leader) base = atomicAdd(A.work, (uint32_t)__popc(need));
                base = __shfl_sync(0xFFFFFFFFu, base, leader);
                if (more && cur == TRACE_IDLE) {
                    const uint32_t idx = base + __popc(need & ((1u << lane) - 1u));
                    if (idx < ntot) {
                        const bool j1 = idx >= n0;
                        const TraceJob& J = A.job[j1 ? 1 : 0];
                        const uint32_t q = j1 ? idx - n0 : idx;
                        slot = J.slot0 + q;
                        // streaming (evict-first) loads and stores for the ray / list records: the L1 is for the tree
                        const D3 o = d3(__ldcs(J.ray[0] + q), __ldcs(J.ray[1] + q), __ldcs(J.ray[2] + q));
                        const D3 dir = d3(__ldcs(J.ray[3] + q), __ldcs(J.ray[4] + q), __ldcs(J.ray[5] + q));
                        best_ub = 3.0e38f;
                        t_occ = -3.0e38f;
                        ncand = 0;
                        double t0 = 0.0;
                        sp = 0;
                        cur = 0;
                        lk = 0;
                        float fmaxo = fmaxf(fmaxf(fabsf((float)o.x), fabsf((float)o.y)), fabsf((float)o.z));
                        if (ray_needs_exact_path(sc, dir)) {
                            // a zero (or subnormal) direction component: the reference's box tests contain 0 * inf = NaN terms
                            // for this ray, so a hit that is "certain" geometrically may still be one the reference never
                            // reaches.  No f32 culling: k_confirm / k_nee_resolve trace it exactly, chain checks included.
                            HNM_FINISH(CAND_OVERFLOW)
                        } else if (fmaxo > sc.far_limit) {
                            // the f32 origin would lose too many bits: advance it to the scene box first
                            double dist;
                            bool h = aabb_intersect_ray(sc.bounds_lo[0], sc.bounds_lo[1], sc.bounds_lo[2], sc.bounds_hi[0], sc.bounds_hi[1],
                                                        sc.bounds_hi[2], o, dir, &dist);
                            if (!h) HNM_FINISH(0u)  // cannot hit anything: every primitive lies inside the scene box
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
#if HNM_TRACE_FMA_SLABS
                        {
                            const float E = 2.384185791015625e-07f;  // 2^-22
                            const float cx = R.ox * R.ix, cy = R.oy * R.iy, cz = R.oz * R.iz;
                            R.nx = -cx - fabsf(cx) * E; R.ny = -cy - fabsf(cy) * E; R.nz = -cz - fabsf(cz) * E;
                            R.fx = -cx + fabsf(cx) * E; R.fy = -cy + fabsf(cy) * E; R.fz = -cz + fabsf(cz) * E;
                        }
#endif
                    }
                }
                more = more && base + (uint32_t)__popc(need) < ntot;
            }
            if (__ballot_sync(0xFFFFFFFFu, cur != TRACE_IDLE) == 0) break;
        }
        // ---- traverse until too few lanes are left -----------------------------------------------------------
        // Majority-vote scheduling.  Every lane is in one of two states: it holds a NODE (cur >= 0) or a LEAF with
        // primitives left (cur < 0).  The warp executes a short phase of whichever kind the majority of lanes needs
        // -- node steps (two f32 slab tests each) or leaf steps (one f32 pre-test each) -- and the minority idles
        // until it becomes the majority.  The loop is warp-uniform (full-mask ballots): with independent thread
        // scheduling a per-lane `while (has_ray)` never reconverges, and a strict while-while loop (all lanes
        // descend to a leaf, then all test it) ran at 4 of 32 lanes per instruction on the incoherent bounces
        // because the longest descent of the warp sets the pace (ncu, round 1).
        for (;;) {
            const unsigned nm = __ballot_sync(0xFFFFFFFFu, cur >= 0);
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

#if HNM_TRACE_FMA_SLABS
                        // entry / exit plane per axis by the sign of the direction, one fma per plane (12 instead of 24
                        // operations per node; a NaN -- axis-parallel ray -- drops out of the min / max: conservative)
                        const bool px = R.ix >= 0.0f, py = R.iy >= 0.0f, pz = R.iz >= 0.0f;
                        // child 0: lo = (m0.x m0.y m0.z) hi = (m0.w m1.x m1.y); child 1: lo = (m1.z m1.w m2.x) hi = (m2.y m2.z m2.w)
                        float tmin0 = fmaxf(fmaxf(__fmaf_rn(px ? m0.x : m0.w, R.ix, R.nx), __fmaf_rn(py ? m0.y : m1.x, R.iy, R.ny)),
                                            __fmaf_rn(pz ? m0.z : m1.y, R.iz, R.nz));
                        float tmax0 = fminf(fminf(__fmaf_rn(px ? m0.w : m0.x, R.ix, R.fx), __fmaf_rn(py ? m1.x : m0.y, R.iy, R.fy)),
                                            __fmaf_rn(pz ? m1.y : m0.z, R.iz, R.fz));
                        float tmin1 = fmaxf(fmaxf(__fmaf_rn(px ? m1.z : m2.y, R.ix, R.nx), __fmaf_rn(py ? m1.w : m2.z, R.iy, R.ny)),
                                            __fmaf_rn(pz ? m2.x : m2.w, R.iz, R.nz));
                        float tmax1 = fminf(fminf(__fmaf_rn(px ? m2.y : m1.z, R.ix, R.fx), __fmaf_rn(py ? m2.z : m1.w, R.iy, R.fy)),
                                            __fmaf_rn(pz ? m2.w : m2.x, R.iz, R.fz));
#else
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
#endif
                        // relative margin for the f32 roundings: widen towards 0 / +inf.  (For a negative bound the product
                        // moves the other way, by 2^-21 relative -- irrelevant: a box with tmax < 0 is behind the ray either
                        // way, and a negative tmin only ever meets the tests `<= up` and `<= best_ub` with non-negative right sides.)
                        float lo0 = tmin0 * WLO, up0 = tmax0 * WUP;
                        float lo1 = tmin1 * WLO, up1 = tmax1 * WUP;
                        const bool h0 = (lo0 <= up0) && (up0 >= 0.0f) && (lo0 <= best_ub);
                        const bool h1 = (lo1 <= up1) && (up1 >= 0.0f) && (lo1 <= best_ub);
                        const bool swap = lo1 < lo0;
                        lk = 0;
                        if (h0 || h1) {
                            // near child next; the far one (if both are hit) goes on the stack
                            cur = (h0 && !(h1 && swap)) ? m3.x : m3.y;
                            if (h0 && h1 && sp < HNM_STACK) stack[sp++] = swap ? m3.x : m3.y;
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
                        bool leaf_done = true, cand = false;
                        uint32_t cid = 0;
                        float tlo = 0.f;
                        if (kind == LEAF_TRI) {
                            const uint32_t pos = first + lk;
                            cand = tri_pretest(sc.trif + 3 * (size_t)pos, R, best_ub, &tlo);
                            if (cand) cid = __ldg(sc.tri_perm + pos);  // index in the reference's leaf order (ties)
                            lk++;
                            leaf_done = lk >= leaf_count(cur);
                        } else if (kind == LEAF_SPHERE) {
                            cand = sphere_pretest(sc.elemf + 4 * (size_t)first, R, best_ub, &tlo);
                            cid = ((uint32_t)LEAF_SPHERE << 30) | first;
                        } else if (kind == LEAF_CUBOID) {
                            cand = cuboid_pretest(sc.elemf + 4 * (size_t)first, R, best_ub, &tlo);
                            cid = ((uint32_t)LEAF_CUBOID << 30) | first;
                        }
                        if (cand) {
                            if (ncand < (uint32_t)TRACE_CAND) {
                                const size_t at = (size_t)ncand * A.cand.stride + slot;
                                __stcs(A.cand.id + at, cid);
                                __stcs(A.cand.lo + at, tlo);
                                ncand++;
                            } else {
                                ncand = CAND_OVERFLOW;  // k_confirm traces this ray exactly
                            }
                        }
                        if (ncand == CAND_OVERFLOW) {
                            HNM_FINISH(CAND_OVERFLOW)
                        } else if (best_ub < t_occ) {
                            // bounded query: something certainly lies in front of the point the caller asked about
                            HNM_FINISH(CAND_OCCLUDED)
                        } else if (leaf_done) {
                            lk = 0;
                            HNM_POP()
                        }
                    }
                }
            }
        }
    }
#undef HNM_FINISH
#undef HNM_POP
    if (blockIdx.x == 0 && threadIdx.x == 0 && A.stat_segments >= 0) atomicAdd(&A.stats[A.stat_segments], (unsigned long long)n0);
    if (STATS) {
        for (int s = 16; s > 0; s >>= 1) n_nodes += __shfl_xor_sync(0xFFFFFFFFu, n_nodes, s);
        if (lane == 0) atomicAdd(&A.stats[A.stat_nodes], (unsigned long long)n_nodes);
    }
}

// The exact closest hit of one ray from its candidate list (or by the plain exact traversal if the list overflowed).
// The header and entry 0 are passed in so that the caller can request them together with everything else it needs.
template <bool STATS>
HNM_D Hit confirm_ray(const DScene& sc, const CandLists& cand, uint32_t slot, uint32_t n, float ub, uint32_t cid0, float lo0, D3 o, D3 dir,
                      uint32_t& n_prims) {
    Hit best;
    best.t = sc.inf; best.u = 0.0; best.v = 0.0; best.kind = LEAF_NONE; best.id = 0;
    if (n == CAND_OVERFLOW) {
        TraceStats st{0, 0};
        best = trace_pretested<STATS>(sc, o, dir, ub, &st);
        if (STATS) n_prims += st.prims;
    } else if (n != CAND_OCCLUDED) {
        for (uint32_t k = 0; k < n; k++) {
            const size_t at = (size_t)k * cand.stride + slot;
            const float lo = k == 0 ? lo0 : __ldcs(cand.lo + at);
            if (lo > ub) continue;  // culled after it was listed
            const uint32_t cid = k == 0 ? cid0 : __ldcs(cand.id + at);
            const uint32_t kind = cid >> 30, id = cid & 0x3FFFFFFFu;
            if (STATS) n_prims++;
            if (kind == LEAF_TRI) {
                DTri tr = load_tri(sc.tris + id);
                tri_test(sc, tr, id, o, dir, best);
            } else if (kind == LEAF_SPHERE) {
                sphere_test(sc, sc.elements[id], id, sc.elements, o, dir, best);
            } else {
                cuboid_test(sc, sc.elements[id], id, sc.elements, o, dir, best);
            }
        }
    }
    return best;
}

// Exact closest hit of every ray from its candidate list; hit records; shading queues for job 0.
// (The NEE shadow rays of the path tracer are confirmed inside k_nee_resolve instead: their hit is consumed there
// and nowhere else.)
template <bool STATS>
__global__ void __launch_bounds__(256, HNM_CONFIRM_MIN_BLOCKS) k_confirm(DScene sc, TraceArgs A) {
    const int lane = threadIdx.x & 31;
    const uint32_t n0 = *A.job[0].count;
    const uint32_t n1 = A.njobs > 1 ? *A.job[1].count : 0u;
    const uint32_t ntot = n0 + n1;
    uint32_t n_prims = 0;
    const bool classify = A.job[0].q_miss != nullptr;
#ifndef HNM_DYN_CONFIRM
#define HNM_DYN_CONFIRM 1
#endif
#if HNM_DYN_CONFIRM
    for (;;) {
        // 32 consecutive rays per fetch (dynamic: see fetch_warp in hnm_kernels.cuh); warp-uniform: the queue pushes ballot
        uint32_t wbase = 0;
        if (lane == 0) wbase = atomicAdd(A.work_confirm, 32u);
        wbase = __shfl_sync(0xFFFFFFFFu, wbase, 0);
        if (wbase >= ntot) break;
        const uint32_t idx = wbase + lane;
#else
    const uint32_t n_round = (ntot + 31u) & ~31u;
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_round; idx += gridDim.x * blockDim.x) {
#endif
        int cls = -1;
        if (idx < ntot) {
            const bool j1 = idx >= n0;
            const TraceJob& J = A.job[j1 ? 1 : 0];
            const uint32_t q = j1 ? idx - n0 : idx;
            const uint32_t slot = J.slot0 + q;
            // everything that does not depend on the list is requested up front: the kernel is a chain of dependent
            // loads (header -> entries -> triangle) and latency is all it costs (ncu, round 1: 15 long-scoreboard
            // stall cycles per issued instruction)
            const uint32_t n = __ldcs(A.cand.n + slot);
            if (STATS && n == CAND_OVERFLOW) atomicAdd(&A.stats[6], 1ull);  // S_OVERFLOW
            const float ub = __ldcs(A.cand.ub + slot);
            const uint32_t cid0 = __ldcs(A.cand.id + slot);   // entry 0 (unspecified if n == 0, always readable)
            const float lo0 = __ldcs(A.cand.lo + slot);
            const D3 o = d3(__ldcs(J.ray[0] + q), __ldcs(J.ray[1] + q), __ldcs(J.ray[2] + q));
            const D3 dir = d3(__ldcs(J.ray[3] + q), __ldcs(J.ray[4] + q), __ldcs(J.ray[5] + q));
            const Hit best = confirm_ray<STATS>(sc, A.cand, slot, n, ub, cid0, lo0, o, dir, n_prims);
            __stcs(J.hit_t + q, best.t); __stcs(J.hit_u + q, best.u); __stcs(J.hit_v + q, best.v);
            __stcs(J.hit_id + q, make_uint2(best.kind, best.id));
            if (!j1 && classify) {
                if (best.kind == LEAF_NONE) cls = 0;
                else {
                    uint32_t el = best.kind == LEAF_TRI ? sc.tri_elem[best.id] : best.id;
                    int surface = sc.materials[sc.elements[el].material].surface;
                    cls = nee_available(surface) ? 2 : 1;
                }
            }
        }
        if (classify) {
            // three queues, ONE round trip: lanes 0..2 each reserve the warp's slots in one queue at the same time
            // (ncu, round 1: the three back-to-back atomics were 14 % of this kernel's stall samples)
            const TraceJob& J0 = A.job[0];
            const unsigned m0 = __ballot_sync(0xFFFFFFFFu, cls == 0), m1 = __ballot_sync(0xFFFFFFFFu, cls == 1),
                           m2 = __ballot_sync(0xFFFFFFFFu, cls == 2);
            uint32_t base = 0;
            if (lane < 3) {
                const unsigned m = lane == 0 ? m0 : (lane == 1 ? m1 : m2);
                uint32_t* ctr = lane == 0 ? J0.cnt_miss : (lane == 1 ? J0.cnt_delta : J0.cnt_nee);
                if (m) base = atomicAdd(ctr, (uint32_t)__popc(m));
            }
            const uint32_t b0 = __shfl_sync(0xFFFFFFFFu, base, 0), b1 = __shfl_sync(0xFFFFFFFFu, base, 1), b2 = __shfl_sync(0xFFFFFFFFu, base, 2);
            if (cls >= 0) {
                const unsigned m = cls == 0 ? m0 : (cls == 1 ? m1 : m2);
                uint32_t* q = cls == 0 ? J0.q_miss : (cls == 1 ? J0.q_delta : J0.q_nee);
                q[(cls == 0 ? b0 : (cls == 1 ? b1 : b2)) + __popc(m & ((1u << lane) - 1u))] = idx;
            }
        }
    }
    if (STATS) {
        for (int s = 16; s > 0; s >>= 1) n_prims += __shfl_xor_sync(0xFFFFFFFFu, n_prims, s);
        if (lane == 0) atomicAdd(&A.stats[A.stat_prims], (unsigned long long)n_prims);
    }
}


// ------------------------------------------------------------------------------------------------ k_confirm, pair-parallel
// k_confirm gives every lane its own ray and walks that ray's list: 45 % of the camera rays have no candidate at all and a
// few have several, so the heavy part -- the f64 triangle test and the reference's box chain, ~400 instructions behind
// three dependent loads -- ran at 4-8 of 32 lanes, once per list POSITION (ncu, 75 k-triangle scene: 8.7 lanes per
// instruction, 11.6 long-scoreboard stall cycles per issued instruction).  Here the (ray, candidate) PAIRS of a warp's 32
// rays are spread over the lanes: a prefix sum of the per-ray counts numbers the pairs, lane j of a round takes pair j (binary
// search for its owner over the prefix sums, five shuffles), loads that ray and that candidate itself, runs the SAME exact
// test into a private Hit, and the owners pull their pairs' results back with shuffles and keep the best.  The closest hit is
// order independent by construction (argmin t, ties by the reference's DFS order: `better`), so the result is the one
// confirm_ray computes.  32 rays are now one round of dependent loads instead of max-list-length rounds.
HNM_D bool hit_better(const DScene& sc, const Hit& h, const Hit& best) {
    if (h.kind == LEAF_NONE) return false;
    if (best.kind == LEAF_NONE) return true;
    if (h.t < best.t) return true;
    if (h.t > best.t) return false;
    // exact tie (tri_test / sphere_test / cuboid_test): a triangle beats an element, the later triangle beats the earlier, the
    // earlier element (top-level DFS order) beats the later
    if (h.kind == LEAF_TRI) return best.kind != LEAF_TRI || h.id > best.id;
    if (best.kind == LEAF_TRI) return false;
    return sc.elements[h.id].seq < sc.elements[best.id].seq;
}
template <bool STATS>
__global__ void __launch_bounds__(256, HNM_CONFIRM_MIN_BLOCKS) k_confirm_pairs(DScene sc, TraceArgs A) {
    const int lane = threadIdx.x & 31;
    const TraceJob& J = A.job[0];
    const uint32_t n0 = *J.count;
    uint32_t n_prims = 0;
    const bool classify = J.q_miss != nullptr;
    const unsigned FULL = 0xFFFFFFFFu;
    for (;;) {
        uint32_t wbase = 0;
        if (lane == 0) wbase = atomicAdd(A.work_confirm, 32u);
        wbase = __shfl_sync(FULL, wbase, 0);
        if (wbase >= n0) break;
        const uint32_t idx = wbase + lane;
        const bool have = idx < n0;
        const uint32_t slot = J.slot0 + idx;
        // ---- this lane's ray: header and the list positions that survive the final bound
        uint32_t n = 0;
        float ub = 0.f;
        if (have) { n = __ldcs(A.cand.n + slot); ub = __ldcs(A.cand.ub + slot); }
        const bool overflow = have && n == CAND_OVERFLOW;
        if (STATS && overflow) atomicAdd(&A.stats[6], 1ull);  // S_OVERFLOW
        const uint32_t listed = (have && n != CAND_OVERFLOW && n != CAND_OCCLUDED) ? n : 0u;
        uint32_t vm = 0;
        const uint32_t maxl = __reduce_max_sync(FULL, listed);
        for (uint32_t k = 0; k < maxl; k++)
            if (k < listed) {
                const float lo = __ldcs(A.cand.lo + (size_t)k * A.cand.stride + slot);
                if (!(lo > ub)) vm |= 1u << k;  // (culled after it was listed otherwise)
            }
        const uint32_t cnt = __popc(vm);
        uint32_t end = cnt;  // inclusive prefix sum
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const uint32_t v = __shfl_up_sync(FULL, end, s);
            if (lane >= s) end += v;
        }
        const uint32_t off = end - cnt;
        const uint32_t T = __shfl_sync(FULL, end, 31);
        const uint32_t maxc = __reduce_max_sync(FULL, cnt);
        if (STATS) n_prims += cnt;
        Hit best;
        best.t = sc.inf; best.u = 0.0; best.v = 0.0; best.kind = LEAF_NONE; best.id = 0;
        for (uint32_t j0 = 0; j0 < T; j0 += 32) {
            const uint32_t j = j0 + lane;
            // owner of pair j: the first lane whose inclusive prefix exceeds j
            int a = 0, b = 31;
#pragma unroll
            for (int it = 0; it < 5; it++) {
                const int mid = (a + b) >> 1;
                const uint32_t e = __shfl_sync(FULL, end, mid);
                if (e > j) b = mid; else a = mid + 1;
            }
            const int L = a & 31;
            const uint32_t offL = __shfl_sync(FULL, off, L), vmL = __shfl_sync(FULL, vm, L);
            Hit h;
            h.t = sc.inf; h.u = 0.0; h.v = 0.0; h.kind = LEAF_NONE; h.id = 0;
            if (j < T) {
                const uint32_t k = __fns(vmL, 0, (int)(j - offL) + 1);  // list position of this pair
                const uint32_t q = wbase + (uint32_t)L;
                const uint32_t cid = __ldcs(A.cand.id + (size_t)k * A.cand.stride + (J.slot0 + q));
                const D3 o = d3(J.ray[0][q], J.ray[1][q], J.ray[2][q]);
                const D3 dir = d3(J.ray[3][q], J.ray[4][q], J.ray[5][q]);
                const uint32_t kind = cid >> 30, id = cid & 0x3FFFFFFFu;
                if (kind == LEAF_TRI) {
                    DTri tr = load_tri(sc.tris + id);
                    tri_test(sc, tr, id, o, dir, h);
                } else if (kind == LEAF_SPHERE) {
                    sphere_test(sc, sc.elements[id], id, sc.elements, o, dir, h);
                } else {
                    cuboid_test(sc, sc.elements[id], id, sc.elements, o, dir, h);
                }
            }
            // the owners pull the results of their pairs of this round
            for (uint32_t r = 0; r < maxc; r++) {
                const int src = (int)(off + r) - (int)j0;
                Hit g;
                g.t = __shfl_sync(FULL, h.t, src & 31);
                g.u = __shfl_sync(FULL, h.u, src & 31);
                g.v = __shfl_sync(FULL, h.v, src & 31);
                g.kind = __shfl_sync(FULL, h.kind, src & 31);
                g.id = __shfl_sync(FULL, h.id, src & 31);
                if (r < cnt && src >= 0 && src < 32 && hit_better(sc, g, best)) best = g;
            }
        }
        int cls = -1;
        if (have) {
            if (overflow) {
                // the list overflowed (or the ray must not be culled in f32 at all): plain exact traversal
                const D3 o = d3(__ldcs(J.ray[0] + idx), __ldcs(J.ray[1] + idx), __ldcs(J.ray[2] + idx));
                const D3 dir = d3(__ldcs(J.ray[3] + idx), __ldcs(J.ray[4] + idx), __ldcs(J.ray[5] + idx));
                TraceStats st{0, 0};
                best = trace_pretested<STATS>(sc, o, dir, ub, &st);
                if (STATS) n_prims += st.prims;
            }
            __stcs(J.hit_t + idx, best.t); __stcs(J.hit_u + idx, best.u); __stcs(J.hit_v + idx, best.v);
            __stcs(J.hit_id + idx, make_uint2(best.kind, best.id));
            if (classify) {
                if (best.kind == LEAF_NONE) cls = 0;
                else {
                    uint32_t el = best.kind == LEAF_TRI ? sc.tri_elem[best.id] : best.id;
                    int surface = sc.materials[sc.elements[el].material].surface;
                    cls = nee_available(surface) ? 2 : 1;
                }
            }
        }
        if (classify) {
            // three queues, ONE round trip: lanes 0..2 each reserve the warp's slots in one queue at the same time
            const unsigned m0 = __ballot_sync(FULL, cls == 0), m1 = __ballot_sync(FULL, cls == 1), m2 = __ballot_sync(FULL, cls == 2);
            uint32_t base = 0;
            if (lane < 3) {
                const unsigned m = lane == 0 ? m0 : (lane == 1 ? m1 : m2);
                uint32_t* ctr = lane == 0 ? J.cnt_miss : (lane == 1 ? J.cnt_delta : J.cnt_nee);
                if (m) base = atomicAdd(ctr, (uint32_t)__popc(m));
            }
            const uint32_t b0 = __shfl_sync(FULL, base, 0), b1 = __shfl_sync(FULL, base, 1), b2 = __shfl_sync(FULL, base, 2);
            if (cls >= 0) {
                const unsigned m = cls == 0 ? m0 : (cls == 1 ? m1 : m2);
                uint32_t* qq = cls == 0 ? J.q_miss : (cls == 1 ? J.q_delta : J.q_nee);
                qq[(cls == 0 ? b0 : (cls == 1 ? b1 : b2)) + __popc(m & ((1u << lane) - 1u))] = idx;
            }
        }
    }
    if (STATS) {
        for (int s = 16; s > 0; s >>= 1) n_prims += __shfl_xor_sync(FULL, n_prims, s);
        if (lane == 0) atomicAdd(&A.stats[A.stat_prims], (unsigned long long)n_prims);
    }
}


// ------------------------------------------------------------------------------------------------ k_confirm, TMA-staged (A/B)
// The same kernel with its streaming inputs -- the six ray arrays and the four list-header arrays of 256 consecutive rays
// -- staged into shared memory by 1-D bulk copies (cp.async.bulk ... mbarrier::complete_tx::bytes), double buffered: one
// thread arms an mbarrier with the byte count and issues ten copies for the NEXT 256 rays while the CTA works on the
// current ones.  `north_star` asks for TMA staging; this is where it applies on this path (contiguous SoA runs; the tree
// itself is reached by per-lane gathers, which a bulk copy cannot express).  HNM_CONFIRM_TMA=1 selects it; DESIGN.md has
// the measured A/B.  It needs 32 KB of shared memory per CTA, which cannot co-reside with the generation kernel's 224 KB.
HNM_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
HNM_D void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
HNM_D void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
HNM_D void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}
HNM_D void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "HNM_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra HNM_DONE_%=;\n"
        "bra HNM_WAIT_%=;\n"
        "HNM_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
template <bool STATS>
__global__ void __launch_bounds__(256, HNM_CONFIRM_MIN_BLOCKS) k_confirm_tma(DScene sc, TraceArgs A) {
    constexpr int CH = 256;
    __shared__ __align__(128) double s_ray[2][6][CH];
    __shared__ __align__(128) uint32_t s_n[2][CH];
    __shared__ __align__(128) float s_ub[2][CH];
    __shared__ __align__(128) uint32_t s_id0[2][CH];
    __shared__ __align__(128) float s_lo0[2][CH];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ uint32_t s_base[2];
    const int lane = threadIdx.x & 31;
    const TraceJob& J = A.job[0];
    const uint32_t ntot = *J.count;
    uint32_t n_prims = 0;
    const bool classify = J.q_miss != nullptr;
    // one thread: claim the next 256 rays and start their ten copies into stage `st`
    auto prefetch = [&](int st) {
        const uint32_t base = atomicAdd(A.work_confirm, (uint32_t)CH);
        s_base[st] = base;
        if (base >= ntot) return;
        const uint32_t cnt = ((ntot - base < (uint32_t)CH ? ntot - base : (uint32_t)CH) + 3u) & ~3u;  // 16-byte multiples (arrays are padded)
        mbar_expect_tx(&s_bar[st], cnt * (6u * 8u + 4u * 4u));
        const uint32_t slot = J.slot0 + base;
        for (int k = 0; k < 6; k++) bulk_g2s(&s_ray[st][k][0], J.ray[k] + base, cnt * 8u, &s_bar[st]);
        bulk_g2s(&s_n[st][0], A.cand.n + slot, cnt * 4u, &s_bar[st]);
        bulk_g2s(&s_ub[st][0], A.cand.ub + slot, cnt * 4u, &s_bar[st]);
        bulk_g2s(&s_id0[st][0], A.cand.id + slot, cnt * 4u, &s_bar[st]);
        bulk_g2s(&s_lo0[st][0], A.cand.lo + slot, cnt * 4u, &s_bar[st]);
    };
    if (threadIdx.x == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        prefetch(0);
    }
    __syncthreads();
    for (uint32_t it = 0;; it++) {
        const int st = it & 1;
        if (threadIdx.x == 0) prefetch(st ^ 1);  // stage st^1 was released by the barrier that ended the previous iteration
        const uint32_t base = s_base[st];
        if (base >= ntot) break;                 // CTA-uniform: s_base[st] was published before the last barrier
        mbar_wait(&s_bar[st], (it >> 1) & 1u);
        const uint32_t idx = base + threadIdx.x;
        int cls = -1;
        if (idx < ntot) {
            const uint32_t q = idx, slot = J.slot0 + q;
            const uint32_t n = s_n[st][threadIdx.x];
            if (STATS && n == CAND_OVERFLOW) atomicAdd(&A.stats[6], 1ull);
            const D3 o = d3(s_ray[st][0][threadIdx.x], s_ray[st][1][threadIdx.x], s_ray[st][2][threadIdx.x]);
            const D3 dir = d3(s_ray[st][3][threadIdx.x], s_ray[st][4][threadIdx.x], s_ray[st][5][threadIdx.x]);
            const Hit best = confirm_ray<STATS>(sc, A.cand, slot, n, s_ub[st][threadIdx.x], s_id0[st][threadIdx.x], s_lo0[st][threadIdx.x], o, dir, n_prims);
            __stcs(J.hit_t + q, best.t); __stcs(J.hit_u + q, best.u); __stcs(J.hit_v + q, best.v);
            __stcs(J.hit_id + q, make_uint2(best.kind, best.id));
            if (classify) {
                if (best.kind == LEAF_NONE) cls = 0;
                else {
                    uint32_t el = best.kind == LEAF_TRI ? sc.tri_elem[best.id] : best.id;
                    int surface = sc.materials[sc.elements[el].material].surface;
                    cls = nee_available(surface) ? 2 : 1;
                }
            }
        }
        if (classify) {
            const unsigned m0 = __ballot_sync(0xFFFFFFFFu, cls == 0), m1 = __ballot_sync(0xFFFFFFFFu, cls == 1), m2 = __ballot_sync(0xFFFFFFFFu, cls == 2);
            uint32_t b = 0;
            if (lane < 3) {
                const unsigned m = lane == 0 ? m0 : (lane == 1 ? m1 : m2);
                uint32_t* ctr = lane == 0 ? J.cnt_miss : (lane == 1 ? J.cnt_delta : J.cnt_nee);
                if (m) b = atomicAdd(ctr, (uint32_t)__popc(m));
            }
            const uint32_t b0 = __shfl_sync(0xFFFFFFFFu, b, 0), b1 = __shfl_sync(0xFFFFFFFFu, b, 1), b2 = __shfl_sync(0xFFFFFFFFu, b, 2);
            if (cls >= 0) {
                const unsigned m = cls == 0 ? m0 : (cls == 1 ? m1 : m2);
                uint32_t* q = cls == 0 ? J.q_miss : (cls == 1 ? J.q_delta : J.q_nee);
                q[(cls == 0 ? b0 : (cls == 1 ? b1 : b2)) + __popc(m & ((1u << lane) - 1u))] = idx;
            }
        }
        __syncthreads();  // stage st may be refilled, s_base[st^1] is visible
    }
    if (STATS) {
        for (int s = 16; s > 0; s >>= 1) n_prims += __shfl_xor_sync(0xFFFFFFFFu, n_prims, s);
        if (lane == 0) atomicAdd(&A.stats[A.stat_prims], (unsigned long long)n_prims);
    }
}

}  // namespace hnm
#endif
