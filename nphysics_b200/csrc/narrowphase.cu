// Device manifold producer and the per-step contact refresh.
//
// The reference takes its ColliderContactManifolds from ncollide's broad and narrow phase
// (src/world/geometrical_world.rs:285-320, call sites src/world/mechanical_world.rs:250-253, 375-385)
// and hands them to MoreauJeanSolver::step by reference.  ncollide is a third-party dependency that
// is not under /root/reference; what is restated here is the producer this repository's own scenes
// use (nphysics_b200/scenes.py ContactGenerator): persistent face-face feature pairs between (near)
// axis-aligned cuboid colliders, each pair yielding one manifold with the <= 4 corners of the overlap
// rectangle as Plane/Point (or Point/Plane) contacts that are re-evaluated from the CURRENT poses every
// step, with stable contact ids -- the records ncollide's polyhedral clipping produces for box piles
// (SURVEY.md appendix B, C, D).
//
//   nb2_upload_colliders   cuboid colliders (ncollide Cuboid + the Collider fields the contact path reads)
//   nb2_detect_pairs       broad phase (uniform hash grid, bodies against their 27 neighbour cells; the few
//                          non-dynamic colliders against every dynamic one) + feature discovery
//   nb2_generate_manifolds one thread per pair: nb2_manifold + <= 4 nb2_contact records written straight
//                          into the buffers nb2_upload_manifolds would have filled -- nothing crosses PCIe
//   nb2_update_contacts    hosts that keep their own narrow phase: only the 40 bytes of a TrackedContact
//                          that change every step (world1, world2, normal, depth) are uploaded
//
// Canonical pair order (deterministic, no sort over all pairs): pairs are owned by a DYNAMIC collider o and
// listed owner by owner: first o's non-dynamic partners g (pair (g, o), ascending g), then its dynamic
// partners b > o (pair (o, b), ascending b).  Contact i of pair p has the id 4 p + i + 1.
#include <stdlib.h>

#include "solver.cuh"

namespace nb2 {

static const int TPB = 128;
static inline unsigned int nblk(size_t n) { return (unsigned int)((n + TPB - 1) / TPB); }

static_assert(sizeof(nb2_collider) == 64, "nb2_collider is read as four quads");
static_assert(sizeof(nb2_contact_update) == 40, "nb2_contact_update layout");

struct ColliderRec {
    Vec3 he;
    float margin;
    Vec3 t;
    float friction;
    Quat r;
    float restitution;
    int body;
    unsigned int modes;  // friction_mode | restitution_mode << 8
};
__device__ __forceinline__ ColliderRec load_collider(const nb2_collider* __restrict__ colliders, int i) {
    const float4* q = reinterpret_cast<const float4*>(&colliders[i]);
    const float4 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2), q3 = __ldg(q + 3);
    ColliderRec c;
    c.he = mk3(q0.x, q0.y, q0.z);
    c.margin = q0.w;
    c.t = mk3(q1.x, q1.y, q1.z);
    c.friction = q1.w;
    c.r = mkq(q2.x, q2.y, q2.z, q2.w);
    c.restitution = q3.x;
    c.body = __float_as_int(q3.y);
    c.modes = (unsigned int)__float_as_int(q3.z) & 0xFFFFu;
    return c;
}

// np_par (device-resident parameters of the producer): [0] max half extent over dynamic colliders (float
// bits), [1] max margin (float bits), [2] search radius (float bits), [3] number of non-dynamic colliders
#define NP_RMAX 0
#define NP_MARGIN 1
#define NP_RADIUS 2
#define NP_NBIG 3

// collider centre in world space + classification
__global__ void k_collider_world(const nb2_collider* __restrict__ colliders, unsigned int n,
                                 ConstPoseQuads pos_t, ConstPoseQuads pos_q,
                                 const int* __restrict__ status, float4* cw, unsigned int* is_big, unsigned int* np_par) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const ColliderRec c = load_collider(colliders, (int)i);
    const Vec3 centre = f4_xyz(pos_t[c.body]) + quat_rotate(f4_quat(pos_q[c.body]), c.t);
    const bool has = fmaxf(c.he.x, fmaxf(c.he.y, c.he.z)) > 0.f;
    const int st = status[c.body];
    const bool dyn = has && (st == NB2_BODY_DYNAMIC || st == NB2_BODY_MULTIBODY_LINK);  // a multibody link collides like a body
    const bool big = has && !dyn && st != NB2_BODY_DISABLED;
    cw[i] = xyz_f4(centre, dyn ? 1.f : 0.f);
    is_big[i] = big ? 1u : 0u;
    if (dyn) atomicMax(&np_par[NP_RMAX], __float_as_uint(fmaxf(c.he.x, fmaxf(c.he.y, c.he.z))));
    if (has) atomicMax(&np_par[NP_MARGIN], __float_as_uint(fmaxf(c.margin, 0.f)));
}
__global__ void k_np_finish_params(unsigned int* np_par, float prediction, float search, const unsigned int* big_off,
                                   unsigned int n) {
    const float rmax = __uint_as_float(np_par[NP_RMAX]), margin = __uint_as_float(np_par[NP_MARGIN]);
    const float reach = 2.f * (margin + prediction);
    // scenes.py ContactGenerator: (2 r_max + reach) * sqrt(2) * 1.01 unless the caller fixes the radius
    const float radius = search >= 0.f ? search : (2.f * rmax + reach) * 1.41421356f * 1.01f;
    np_par[NP_RADIUS] = __float_as_uint(radius);
    np_par[NP_NBIG] = big_off[n];
}
__global__ void k_fill_big(const unsigned int* __restrict__ is_big, const unsigned int* __restrict__ big_off,
                           unsigned int n, int* big_list) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && is_big[i]) big_list[big_off[i]] = (int)i;
}

__device__ __forceinline__ int3 cell_of(Vec3 c, float inv_cell) {
    return make_int3((int)floorf(c.x * inv_cell), (int)floorf(c.y * inv_cell), (int)floorf(c.z * inv_cell));
}
__device__ __forceinline__ unsigned int cell_hash(int3 c, unsigned int mask) {
    return (((unsigned int)c.x * 73856093u) ^ ((unsigned int)c.y * 19349663u) ^ ((unsigned int)c.z * 83492791u)) & mask;
}
__device__ __forceinline__ float cell_size(const unsigned int* np_par) {
    const float r = __uint_as_float(np_par[NP_RADIUS]);
    return r > 1e-6f ? r : 1.f;
}
__global__ void k_grid_count(const float4* __restrict__ cw, unsigned int n, const unsigned int* __restrict__ np_par,
                             unsigned int mask, unsigned int* count) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 c = cw[i];
    if (c.w == 0.f) return;
    atomicAdd(&count[cell_hash(cell_of(f4_xyz(c), 1.f / cell_size(np_par)), mask)], 1u);
}
__global__ void k_grid_fill(const float4* __restrict__ cw, unsigned int n, const unsigned int* __restrict__ np_par,
                            unsigned int mask, const unsigned int* __restrict__ off, unsigned int* cursor, int* entries) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 c = cw[i];
    if (c.w == 0.f) return;
    const unsigned int h = cell_hash(cell_of(f4_xyz(c), 1.f / cell_size(np_par)), mask);
    entries[off[h] + atomicAdd(&cursor[h], 1u)] = (int)i;
}

// Face-face feature test of scenes.py ContactGenerator.__init__: the first axis along which the two boxes
// face each other within `reach` (and not deeper than half the thinner box) while their projections on the
// other two axes overlap.
__device__ __forceinline__ bool face_axis(Vec3 ca, Vec3 ha, Vec3 cb, Vec3 hb, float reach, int* axis_out) {
    const float eps = 1e-6f;
    const float d[3] = {cb.x - ca.x, cb.y - ca.y, cb.z - ca.z};
    const float a[3] = {ha.x, ha.y, ha.z}, b[3] = {hb.x, hb.y, hb.z};
    float gap[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) gap[k] = fabsf(d[k]) - (a[k] + b[k]);
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
        const int o1 = (ax + 1) % 3, o2 = (ax + 2) % 3;
        if (gap[ax] <= reach + eps && gap[ax] > -0.5f * fminf(a[ax], b[ax]) && -gap[o1] > eps && -gap[o2] > eps) {
            *axis_out = ax;
            return true;
        }
    }
    return false;
}

#define NB2_MAX_PARTNERS 96

struct PairOut {
    float4* q;  // [3][cap]: rectangle rel. a | rectangle rel. b | (face a, face b, a, b)
    int* feat;  // axis | 4 * (sign < 0) | 8 * flipped
    size_t cap;
};

// One thread per collider.  FILL = false: counts the pairs it owns; FILL = true: writes them at pair_off[o].
template <bool FILL>
__global__ void __launch_bounds__(TPB) k_pairs(const nb2_collider* __restrict__ colliders, unsigned int n,
                                               const float4* __restrict__ cw, const unsigned int* __restrict__ np_par,
                                               const int* __restrict__ big_list, unsigned int mask,
                                               const unsigned int* __restrict__ grid_off, const int* __restrict__ entries,
                                               float prediction, unsigned int flip_permille, unsigned int* pair_cnt,
                                               const unsigned int* __restrict__ pair_off, PairOut out, unsigned int* flags) {
    const unsigned int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n) return;
    const float4 co4 = cw[o];
    if (co4.w == 0.f) {
        if (!FILL) pair_cnt[o] = 0u;
        return;
    }
    const ColliderRec ro = load_collider(colliders, (int)o);
    const Vec3 co = f4_xyz(co4);
    int list[NB2_MAX_PARTNERS];
    int cnt = 0;
    bool overflow = false;
    // ---- non-dynamic partners (ground, kinematic platforms): every one of them, in index order
    const unsigned int n_big = np_par[NP_NBIG];
    for (unsigned int gi = 0; gi < n_big; ++gi) {
        const int g = __ldg(&big_list[gi]);
        const ColliderRec rg = load_collider(colliders, g);
        int ax;
        if (face_axis(f4_xyz(__ldg(&cw[g])), rg.he, co, ro.he, (rg.margin + ro.margin) + 2.f * prediction, &ax)) {
            if (cnt < NB2_MAX_PARTNERS) list[cnt++] = g; else overflow = true;
        }
    }
    const int n_bigp = cnt;
    // ---- dynamic partners with a higher index, through the hash grid
    const float radius = __uint_as_float(np_par[NP_RADIUS]);
    const float inv_cell = 1.f / cell_size(np_par);
    const int3 c0 = cell_of(co, inv_cell);
    for (int dz = -1; dz <= 1; ++dz)
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                const int3 cc = make_int3(c0.x + dx, c0.y + dy, c0.z + dz);
                const unsigned int h = cell_hash(cc, mask);
                const unsigned int e0 = grid_off[h], e1 = grid_off[h + 1];
                for (unsigned int e = e0; e < e1; ++e) {
                    const int b = __ldg(&entries[e]);
                    if (b <= (int)o) continue;
                    const Vec3 cb = f4_xyz(__ldg(&cw[b]));
                    const int3 cellb = cell_of(cb, inv_cell);
                    if (cellb.x != cc.x || cellb.y != cc.y || cellb.z != cc.z) continue;  // another cell of this bucket
                    if (norm_sq3(cb - co) > radius * radius) continue;
                    const ColliderRec rb = load_collider(colliders, b);
                    int ax;
                    if (!face_axis(co, ro.he, cb, rb.he, (ro.margin + rb.margin) + 2.f * prediction, &ax)) continue;
                    if (cnt < NB2_MAX_PARTNERS) list[cnt++] = b; else overflow = true;
                }
            }
    if (overflow) atomicOr(flags, 4u);
    if constexpr (!FILL) {
        pair_cnt[o] = (unsigned int)cnt;
    } else {
    for (int i = n_bigp + 1; i < cnt; ++i) {  // bucket order is arbitrary: sort the dynamic partners by index
        const int v = list[i];
        int j = i;
        while (j > n_bigp && list[j - 1] > v) {
            list[j] = list[j - 1];
            --j;
        }
        list[j] = v;
    }
    const size_t base = pair_off[o];
    for (int i = 0; i < cnt; ++i) {
        const int a = i < n_bigp ? list[i] : (int)o, b = i < n_bigp ? (int)o : list[i];
        const ColliderRec ra = i < n_bigp ? load_collider(colliders, a) : ro;
        const ColliderRec rb = i < n_bigp ? ro : load_collider(colliders, b);
        const Vec3 ca = i < n_bigp ? f4_xyz(__ldg(&cw[a])) : co, cb = i < n_bigp ? co : f4_xyz(__ldg(&cw[b]));
        int ax = 0;
        face_axis(ca, ra.he, cb, rb.he, (ra.margin + rb.margin) + 2.f * prediction, &ax);
        const float d[3] = {cb.x - ca.x, cb.y - ca.y, cb.z - ca.z};
        const float ha[3] = {ra.he.x, ra.he.y, ra.he.z}, hb[3] = {rb.he.x, rb.he.y, rb.he.z};
        const int u = (ax + 1) % 3, v = (ax + 2) % 3;
        const float sign = d[ax] >= 0.f ? 1.f : -1.f;
        // overlap rectangle of the two faces, relative to a's centre (and to b's): the local contact points
        const float lo_u = fmaxf(-ha[u], d[u] - hb[u]), hi_u = fminf(ha[u], d[u] + hb[u]);
        const float lo_v = fmaxf(-ha[v], d[v] - hb[v]), hi_v = fminf(ha[v], d[v] + hb[v]);
        const size_t p = base + (size_t)i;
        // ContactGenerator's flipped pairs (body1 carries the Point, body2 the Plane): a hash of the pair
        const unsigned long long hsh = ((unsigned long long)a * 2654435761ull + (unsigned long long)b * 40503ull) % 1000ull;
        const bool flip = hsh < (unsigned long long)flip_permille;
        out.q[0 * out.cap + p] = make_float4(lo_u, hi_u, lo_v, hi_v);
        out.q[1 * out.cap + p] = make_float4(lo_u - d[u], hi_u - d[u], lo_v - d[v], hi_v - d[v]);
        out.q[2 * out.cap + p] = make_float4(sign * ha[ax], -sign * hb[ax], __int_as_float(a), __int_as_float(b));
        out.feat[p] = ax | (sign < 0.f ? 4 : 0) | (flip ? 8 : 0);
    }
    }
}

// MaterialCombineMode::combine (src/material/material.rs:72-86)
__device__ __forceinline__ float combine_coeff_dev(float a, unsigned int ma, float b, unsigned int mb) {
    if (ma == 3u || mb == 3u) return fmaxf(a, b);
    if (ma == 2u || mb == 2u) return a * b;
    if (ma == 1u || mb == 1u) return fminf(a, b);
    return (a + b) * 0.5f;
}

// shared-memory stride of a pair's four contact records in quads: 28 used, 29 keeps a warp's 16-byte stores
// (lanes 29 quads apart) off each other's banks
#define GM_CQ 29
static_assert(sizeof(nb2_manifold) % 4 == 0 && sizeof(nb2_contact) == 112, "producer staging copies whole words / 7 quads per contact");
// One thread per persistent pair: the manifold of the pair at the current poses.  Manifold p owns the
// contact slots [4 p, 4 p + 4); its kept contacts are compacted to the front of that range (corner order)
// and num_contacts says how many there are -- possibly none, in which case the manifold emits no rows.
__global__ void __launch_bounds__(TPB) k_generate_manifolds(unsigned int n_pairs, PairOut pairs,
                                                            const nb2_collider* __restrict__ colliders,
                                                            ConstPoseQuads pos_t,
                                                            ConstPoseQuads pos_q, float prediction,
                                                            nb2_manifold* manifolds, nb2_contact* contacts) {
    // Records are staged in shared memory and leave the block as two contiguous runs (the block's 4 x TPB contact
    // slots, its TPB manifolds): written straight from the threads, every 16-byte store of a warp went to 32
    // different sectors and the kernel sat in the store queue (lg_throttle 15 per issue, 107 us for 296k pairs).
    extern __shared__ __align__(16) unsigned char gm_smem[];
    float4* s_c = reinterpret_cast<float4*>(gm_smem);                                        // [TPB][GM_CQ]: 28 quads used
    nb2_manifold* s_m = reinterpret_cast<nb2_manifold*>(gm_smem + (size_t)TPB * GM_CQ * 16);  // [TPB]
    const unsigned int p0 = blockIdx.x * blockDim.x, p = p0 + threadIdx.x;
    if (p < n_pairs) {
    const float4 ra = pairs.q[0 * pairs.cap + p], rb = pairs.q[1 * pairs.cap + p], fq = pairs.q[2 * pairs.cap + p];
    const int feat = pairs.feat[p];
    const int ax = feat & 3, u = (ax + 1) % 3, v = (ax + 2) % 3;
    const bool flip = (feat & 8) != 0;
    const int a = __float_as_int(fq.z), b = __float_as_int(fq.w);
    const ColliderRec ca = load_collider(colliders, a), cb = load_collider(colliders, b);
    Pose pa, pb, ka, kb;
    pa.t = f4_xyz(pos_t[ca.body]);
    pa.r = f4_quat(pos_q[ca.body]);
    pb.t = f4_xyz(pos_t[cb.body]);
    pb.r = f4_quat(pos_q[cb.body]);
    ka.t = ca.t;
    ka.r = ca.r;
    kb.t = cb.t;
    kb.r = cb.r;
    const Pose wa = pose_mul(pa, ka), wb = pose_mul(pb, kb);  // collider poses in world space
    float na_[3] = {0.f, 0.f, 0.f};
    na_[ax] = (feat & 4) ? -1.f : 1.f;
    const Vec3 na = mk3(na_[0], na_[1], na_[2]);  // outward normal of a's face, local to a
    const Vec3 n = quat_rotate(wa.r, na);         // world normal a -> b
    const float reach = (ca.margin + cb.margin) + 2.f * prediction;
    const float cu_a[4] = {ra.x, ra.y, ra.y, ra.x}, cv_a[4] = {ra.z, ra.z, ra.w, ra.w};
    const float cu_b[4] = {rb.x, rb.y, rb.y, rb.x}, cv_b[4] = {rb.z, rb.z, rb.w, rb.w};
    int kept = 0;
    float4* cq = s_c + (size_t)threadIdx.x * GM_CQ;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float la_[3], lb_[3];
        la_[u] = cu_a[k]; la_[v] = cv_a[k]; la_[ax] = fq.x;
        lb_[u] = cu_b[k]; lb_[v] = cv_b[k]; lb_[ax] = fq.y;
        const Vec3 la = mk3(la_[0], la_[1], la_[2]), lb = mk3(lb_[0], lb_[1], lb_[2]);
        const Vec3 pw_a = pose_point(wa, la), pw_b = pose_point(wb, lb);  // points on a's face / on b's face
        const float depth = -dot3(n, pw_b - pw_a);                           // Plane(a) / Point(b)
        if (!(depth > -reach)) continue;
        const Vec3 w1 = pw_b + n * depth;  // projection of b's point on a's plane
        const Vec3 world1 = flip ? pw_b : w1, world2 = flip ? w1 : pw_b;
        const Vec3 nn = flip ? -n : n;
        const Vec3 l1 = flip ? lb : la, l2 = flip ? la : lb;
        const Vec3 d1 = flip ? mk3(0.f, 0.f, 0.f) : na, d2 = flip ? na : mk3(0.f, 0.f, 0.f);
        const unsigned long long key = (unsigned long long)p * 4ull + (unsigned long long)k + 1ull;
        const unsigned int geoms = flip ? ((unsigned int)NB2_GEOM_POINT | ((unsigned int)NB2_GEOM_PLANE << 8))
                                        : ((unsigned int)NB2_GEOM_PLANE | ((unsigned int)NB2_GEOM_POINT << 8));
        float4* o = cq + 7 * kept;
        o[0] = make_float4(world1.x, world1.y, world1.z, world2.x);
        o[1] = make_float4(world2.y, world2.z, nn.x, nn.y);
        o[2] = make_float4(nn.z, depth, __uint_as_float((unsigned int)(key & 0xFFFFFFFFull)), __uint_as_float((unsigned int)(key >> 32)));
        o[3] = make_float4(l1.x, l1.y, l1.z, l2.x);
        o[4] = make_float4(l2.y, l2.z, d1.x, d1.y);
        o[5] = make_float4(d1.z, d2.x, d2.y, d2.z);
        o[6] = make_float4(0.f, 0.f, __uint_as_float(geoms), 0.f);
        ++kept;
    }
    for (int k = kept; k < 4; ++k) {  // unused slots: null ids, nothing stale for the impulse cache to pick up
        float4* o = cq + 7 * k;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 7; ++j) o[j] = z;
    }
    nb2_manifold& m = s_m[threadIdx.x];
    const ColliderRec& c1 = flip ? cb : ca;
    const ColliderRec& c2 = flip ? ca : cb;
    m.body1 = c1.body;
    m.body2 = c2.body;
    m.first_contact = 4u * p;
    m.num_contacts = (unsigned int)kept;
    m.margin1 = c1.margin;
    m.margin2 = c2.margin;
    m.friction = combine_coeff_dev(ca.friction, ca.modes & 0xFFu, cb.friction, cb.modes & 0xFFu);
    m.restitution = combine_coeff_dev(ca.restitution, (ca.modes >> 8) & 0xFFu, cb.restitution, (cb.modes >> 8) & 0xFFu);
    m.surface_velocity[0] = m.surface_velocity[1] = m.surface_velocity[2] = 0.f;
    m.coll1_wrt_body[0] = c1.t.x; m.coll1_wrt_body[1] = c1.t.y; m.coll1_wrt_body[2] = c1.t.z;
    m.coll1_wrt_body[3] = c1.r.i; m.coll1_wrt_body[4] = c1.r.j; m.coll1_wrt_body[5] = c1.r.k; m.coll1_wrt_body[6] = c1.r.w;
    m.coll2_wrt_body[0] = c2.t.x; m.coll2_wrt_body[1] = c2.t.y; m.coll2_wrt_body[2] = c2.t.z;
    m.coll2_wrt_body[3] = c2.r.i; m.coll2_wrt_body[4] = c2.r.j; m.coll2_wrt_body[5] = c2.r.k; m.coll2_wrt_body[6] = c2.r.w;
    }
    __syncthreads();
    const unsigned int nv = min((unsigned int)TPB, n_pairs - p0);
    float4* gc = reinterpret_cast<float4*>(contacts + (size_t)4 * p0);
    for (unsigned int g = threadIdx.x; g < 28u * nv; g += TPB) gc[g] = s_c[(g / 28u) * GM_CQ + g % 28u];
    unsigned int* gm = reinterpret_cast<unsigned int*>(manifolds + p0);
    const unsigned int* sm = reinterpret_cast<const unsigned int*>(s_m);
    const unsigned int mw = (unsigned int)(sizeof(nb2_manifold) / 4);
    for (unsigned int g = threadIdx.x; g < mw * nv; g += TPB) gm[g] = sm[g];
}

// nb2_update_contacts: the ten floats of a TrackedContact that change from step to step
__global__ void k_apply_contact_updates(const nb2_contact_update* __restrict__ upd, unsigned int n, nb2_contact* contacts) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float2* s = reinterpret_cast<const float2*>(&upd[i]);  // 40 bytes = 5 x 8
    float2* d = reinterpret_cast<float2*>(&contacts[i]);         // world1, world2, normal, depth lead the record
#pragma unroll
    for (int k = 0; k < 5; ++k) d[k] = __ldg(s + k);
}

int launch_upload_colliders(Context* ctx, const nb2_collider* colliders, uint32_t n) {
    NB2_TRY(ctx->colliders.reserve(ctx, n));
    NB2_CUDA(ctx, cudaMemcpyAsync(ctx->colliders.p, colliders, (size_t)n * sizeof(nb2_collider), cudaMemcpyHostToDevice, ctx->stream));
    ctx->n_colliders = n;
    ctx->n_pairs = 0;
    ctx->pairs_valid = false;
    NB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the caller may free `colliders` on return
    return NB2_OK;
}

int launch_detect_pairs(Context* ctx, float prediction, float search, unsigned int flip_permille, uint32_t* out_pairs) {
    const unsigned int n = ctx->n_colliders;
    NB2_TRY(ctx->flags.reserve(ctx, 4));
    NB2_TRY(ctx->coll_world.reserve(ctx, n));
    NB2_TRY(ctx->np_is_big.reserve(ctx, (size_t)n + 1));
    NB2_TRY(ctx->np_big_off.reserve(ctx, (size_t)n + 2));
    NB2_TRY(ctx->np_big_list.reserve(ctx, (size_t)n + 1));
    NB2_TRY(ctx->np_par.reserve(ctx, 8));
    NB2_TRY(ctx->pair_cnt.reserve(ctx, (size_t)n + 1));
    NB2_TRY(ctx->pair_off.reserve(ctx, (size_t)n + 2));
    size_t cells = 64;
    while (cells < 2 * (size_t)n) cells <<= 1;
    NB2_TRY(ctx->grid_count.reserve(ctx, cells + 1));
    NB2_TRY(ctx->grid_off.reserve(ctx, cells + 2));
    NB2_TRY(ctx->grid_cursor.reserve(ctx, cells + 1));
    NB2_TRY(ctx->grid_entries.reserve(ctx, (size_t)n + 1));
    const unsigned int mask = (unsigned int)cells - 1u;
    NB2_CUDA(ctx, cudaMemsetAsync(ctx->np_par.p, 0, 8 * sizeof(unsigned int), ctx->stream));
    NB2_CUDA(ctx, cudaMemsetAsync(ctx->grid_count.p, 0, (cells + 1) * sizeof(unsigned int), ctx->stream));
    NB2_CUDA(ctx, cudaMemsetAsync(ctx->grid_cursor.p, 0, (cells + 1) * sizeof(unsigned int), ctx->stream));
    // classification by the status as uploaded: a sleeping body keeps its pairs
    k_collider_world<<<nblk(n), TPB, 0, ctx->stream>>>(ctx->colliders.p, n, ctx->pos_t.p, ctx->pos_q.p, ctx->true_status.p,
                                                       ctx->coll_world.p, ctx->np_is_big.p, ctx->np_par.p);
    ctx->launches++;
    NB2_TRY(exclusive_scan_u32(ctx, ctx->np_is_big.p, ctx->np_big_off.p, n));
    k_fill_big<<<nblk(n), TPB, 0, ctx->stream>>>(ctx->np_is_big.p, ctx->np_big_off.p, n, ctx->np_big_list.p);
    k_np_finish_params<<<1, 1, 0, ctx->stream>>>(ctx->np_par.p, prediction, search, ctx->np_big_off.p, n);
    k_grid_count<<<nblk(n), TPB, 0, ctx->stream>>>(ctx->coll_world.p, n, ctx->np_par.p, mask, ctx->grid_count.p);
    ctx->launches += 3;
    NB2_TRY(exclusive_scan_u32(ctx, ctx->grid_count.p, ctx->grid_off.p, cells));
    k_grid_fill<<<nblk(n), TPB, 0, ctx->stream>>>(ctx->coll_world.p, n, ctx->np_par.p, mask, ctx->grid_off.p,
                                                  ctx->grid_cursor.p, ctx->grid_entries.p);
    PairOut none;
    none.q = nullptr;
    none.feat = nullptr;
    none.cap = 0;
    k_pairs<false><<<nblk(n), TPB, 0, ctx->stream>>>(ctx->colliders.p, n, ctx->coll_world.p, ctx->np_par.p, ctx->np_big_list.p,
                                                     mask, ctx->grid_off.p, ctx->grid_entries.p, prediction, flip_permille,
                                                     ctx->pair_cnt.p, nullptr, none, ctx->flags.p);
    ctx->launches += 2;
    NB2_TRY(exclusive_scan_u32(ctx, ctx->pair_cnt.p, ctx->pair_off.p, n));
    unsigned int total = 0;
    NB2_CUDA(ctx, cudaMemcpyAsync(&total, ctx->pair_off.p + n, sizeof(total), cudaMemcpyDeviceToHost, ctx->stream));
    NB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the pair count sizes every buffer downstream
    NB2_TRY(ctx->pair_q.reserve(ctx, 3 * ((size_t)total + 16)));
    NB2_TRY(ctx->pair_feat.reserve(ctx, (size_t)total + 16));
    PairOut out;
    out.q = ctx->pair_q.p;
    out.feat = ctx->pair_feat.p;
    out.cap = ctx->pair_q.cap / 3;
    if (total) {
        k_pairs<true><<<nblk(n), TPB, 0, ctx->stream>>>(ctx->colliders.p, n, ctx->coll_world.p, ctx->np_par.p,
                                                        ctx->np_big_list.p, mask, ctx->grid_off.p, ctx->grid_entries.p,
                                                        prediction, flip_permille, ctx->pair_cnt.p, ctx->pair_off.p, out,
                                                        ctx->flags.p);
        ctx->launches++;
    }
    NB2_CUDA(ctx, cudaGetLastError());
    ctx->n_pairs = total;
    ctx->np_prediction = prediction;
    ctx->pairs_valid = true;
    if (out_pairs) *out_pairs = total;
    return NB2_OK;
}

int launch_generate_manifolds(Context* ctx) {
    const unsigned int np = ctx->n_pairs;
    if ((size_t)np > ctx->manifolds.cap || (size_t)4 * np > ctx->contacts.cap)
        NB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // about to reallocate buffers a running step may read
    NB2_TRY(ctx->manifolds.reserve(ctx, np));
    NB2_TRY(ctx->contacts.reserve(ctx, (size_t)4 * np));
    ctx->n_manifolds = np;
    ctx->n_contacts = 4 * np;
    ctx->manifolds_from_producer = true;
    if (np) {
        PairOut pairs;
        pairs.q = ctx->pair_q.p;
        pairs.feat = ctx->pair_feat.p;
        pairs.cap = ctx->pair_q.cap / 3;
        const size_t smem = (size_t)TPB * GM_CQ * 16 + (size_t)TPB * sizeof(nb2_manifold);
        if (!ctx->producer_attr) {
            NB2_CUDA(ctx, cudaFuncSetAttribute(k_generate_manifolds, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ctx->producer_attr = true;
        }
        k_generate_manifolds<<<nblk(np), TPB, smem, ctx->stream>>>(np, pairs, ctx->colliders.p, ctx->pos_t.p, ctx->pos_q.p,
                                                                ctx->np_prediction, ctx->manifolds.p, ctx->contacts.p);
        ctx->launches++;
    }
    NB2_CUDA(ctx, cudaGetLastError());
    return NB2_OK;
}

int launch_update_contacts(Context* ctx, const nb2_contact_update* updates, uint32_t n) {
    NB2_TRY(ctx->contact_updates.reserve(ctx, n));
    NB2_CUDA(ctx, cudaMemcpyAsync(ctx->contact_updates.p, updates, (size_t)n * sizeof(nb2_contact_update), cudaMemcpyHostToDevice,
                                  ctx->stream));
    k_apply_contact_updates<<<nblk(n), TPB, 0, ctx->stream>>>(ctx->contact_updates.p, n, ctx->contacts.p);
    ctx->launches++;
    NB2_CUDA(ctx, cudaGetLastError());
    return NB2_OK;
}

}  // namespace nb2
