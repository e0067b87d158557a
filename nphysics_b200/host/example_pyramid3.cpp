// examples3d/pyramid3.rs over the C++ host mirror: the 465-box pyramid, 8 velocity + 3 position
// iterations, stepped through MechanicalWorld::step.  The manifolds (ncollide's job in the
// reference) come from a small face-face producer for axis-aligned cubes with persistent features.
//
// Build:  g++ -O2 -std=c++17 example_pyramid3.cpp -o example_pyramid3 -L.. -lnphysics_b200 -Wl,-rpath,'$ORIGIN/..'
// Usage:  ./example_pyramid3 [steps] [reference|coloured]
#include <cstdio>
#include <cstdlib>

#include "nphysics_b200.hpp"

using namespace nphysics;

static Vector3 rotate(const Quaternion& q, const Vector3& p) {
    const float vx = q[0], vy = q[1], vz = q[2], w = q[3];
    const float tx = 2.f * (vy * p[2] - vz * p[1]), ty = 2.f * (vz * p[0] - vx * p[2]), tz = 2.f * (vx * p[1] - vy * p[0]);
    return {p[0] + w * tx + (vy * tz - vz * ty), p[1] + w * ty + (vz * tx - vx * tz), p[2] + w * tz + (vx * ty - vy * tx)};
}

struct FacePair {  // Plane (on a) / Point (on b) features of one touching face pair
    size_t a, b;
    Vector3 normal_a;          // outward normal of a's face, local
    Vector3 local_a[4], local_b[4];
    Vector3 offset_a;          // collider offset of a in its body frame
};

class BoxPileContacts {
    std::vector<FacePair> pairs_;
    float margin_, reach_;

  public:
    BoxPileContacts(const DefaultBodySet& bodies, const std::vector<Vector3>& half, const std::vector<Vector3>& offset,
                    float margin)
        : margin_(margin), reach_(2.f * (margin + 0.001f)) {
        const size_t n = bodies.len();
        for (size_t a = 0; a < n; ++a)
            for (size_t b = a + 1; b < n; ++b) {
                const bool da = bodies.get(a)->is_dynamic(), db = bodies.get(b)->is_dynamic();
                if (!da && !db) continue;
                Vector3 ca = bodies.get(a)->position().translation, cb = bodies.get(b)->position().translation;
                for (int k = 0; k < 3; ++k) { ca[k] += offset[a][k]; cb[k] += offset[b][k]; }
                for (int ax = 0; ax < 3; ++ax) {
                    const int u = (ax + 1) % 3, v = (ax + 2) % 3;
                    const float gap = std::fabs(cb[ax] - ca[ax]) - (half[a][ax] + half[b][ax]);
                    const float ou = (half[a][u] + half[b][u]) - std::fabs(cb[u] - ca[u]);
                    const float ov = (half[a][v] + half[b][v]) - std::fabs(cb[v] - ca[v]);
                    if (gap > reach_ + 1e-6f || gap < -0.5f * std::fmin(half[a][ax], half[b][ax]) || ou <= 1e-6f || ov <= 1e-6f)
                        continue;
                    FacePair fp;
                    fp.a = a;
                    fp.b = b;
                    fp.offset_a = offset[a];
                    const float s = cb[ax] >= ca[ax] ? 1.f : -1.f;
                    fp.normal_a = {0.f, 0.f, 0.f};
                    fp.normal_a[ax] = s;
                    const float lo_u = std::fmax(ca[u] - half[a][u], cb[u] - half[b][u]), hi_u = std::fmin(ca[u] + half[a][u], cb[u] + half[b][u]);
                    const float lo_v = std::fmax(ca[v] - half[a][v], cb[v] - half[b][v]), hi_v = std::fmin(ca[v] + half[a][v], cb[v] + half[b][v]);
                    const float cu[4] = {lo_u, hi_u, hi_u, lo_u}, cv[4] = {lo_v, lo_v, hi_v, hi_v};
                    for (int k = 0; k < 4; ++k) {
                        fp.local_a[k][u] = cu[k] - ca[u]; fp.local_a[k][v] = cv[k] - ca[v]; fp.local_a[k][ax] = s * half[a][ax];
                        fp.local_b[k][u] = cu[k] - cb[u]; fp.local_b[k][v] = cv[k] - cb[v]; fp.local_b[k][ax] = -s * half[b][ax];
                    }
                    pairs_.push_back(fp);
                    break;
                }
            }
    }
    std::vector<ColliderContactManifold> generate(const DefaultBodySet& bodies, const std::vector<Vector3>& offset) const {
        std::vector<ColliderContactManifold> out;
        for (size_t pi = 0; pi < pairs_.size(); ++pi) {
            const FacePair& fp = pairs_[pi];
            const Isometry3 pa = bodies.get(fp.a)->position(), pb = bodies.get(fp.b)->position();
            const Vector3 oa = rotate(pa.rotation, offset[fp.a]), ob = rotate(pb.rotation, offset[fp.b]);
            const Vector3 n = rotate(pa.rotation, fp.normal_a);
            ColliderContactManifold m;
            std::memset(&m.manifold, 0, sizeof(m.manifold));
            m.manifold.body1 = (int32_t)fp.a;
            m.manifold.body2 = (int32_t)fp.b;
            m.manifold.margin1 = m.manifold.margin2 = margin_;
            // BasicMaterial::default() on both sides: friction 0.5, restitution 0, Average/Average
            nb2_combine_materials(0.5f, 0, 0.f, 0, nullptr, 0.5f, 0, 0.f, 0, nullptr, &m.manifold.friction, &m.manifold.restitution,
                                  m.manifold.surface_velocity);
            m.manifold.coll1_wrt_body[6] = m.manifold.coll2_wrt_body[6] = 1.f;
            for (int k = 0; k < 3; ++k) { m.manifold.coll1_wrt_body[k] = offset[fp.a][k]; m.manifold.coll2_wrt_body[k] = offset[fp.b][k]; }
            for (int k = 0; k < 4; ++k) {
                const Vector3 ra = rotate(pa.rotation, fp.local_a[k]), rb = rotate(pb.rotation, fp.local_b[k]);
                Vector3 wa, wb;
                for (int d = 0; d < 3; ++d) { wa[d] = pa.translation[d] + oa[d] + ra[d]; wb[d] = pb.translation[d] + ob[d] + rb[d]; }
                const float depth = -(n[0] * (wb[0] - wa[0]) + n[1] * (wb[1] - wa[1]) + n[2] * (wb[2] - wa[2]));
                if (!(depth > -reach_)) continue;
                nb2_contact c;
                std::memset(&c, 0, sizeof(c));
                for (int d = 0; d < 3; ++d) {
                    c.world1[d] = wb[d] + n[d] * depth;
                    c.world2[d] = wb[d];
                    c.normal[d] = n[d];
                    c.local1[d] = fp.local_a[k][d];
                    c.local2[d] = fp.local_b[k][d];
                    c.dir1[d] = fp.normal_a[d];
                }
                c.depth = depth;
                c.key = (uint64_t)pi * 4 + (uint64_t)k + 1;
                c.geom1 = NB2_GEOM_PLANE;
                c.geom2 = NB2_GEOM_POINT;
                m.contacts.push_back(c);
            }
            if (!m.contacts.empty()) out.push_back(m);
        }
        return out;
    }
};

int main(int argc, char** argv) {
    const int steps = argc > 1 ? atoi(argv[1]) : 60;
    const bool reference = argc > 2 && std::string(argv[2]) == "reference";
    try {
        MechanicalWorld world({0.f, -9.81f, 0.f});
        world.solver.mode = reference ? SolverMode::ReferenceOrder : SolverMode::Coloured;
        world.counters.enable();
        DefaultBodySet bodies;
        DefaultJointConstraintSet joints;
        std::vector<Vector3> half, offset;
        // pyramid3.rs:33-41: the ground
        bodies.insert(Ground::make());
        half.push_back({6.f, 0.2f, 6.f});
        offset.push_back({0.f, -0.2f, 0.f});
        // pyramid3.rs:46-72: the boxes
        const int num = 30;
        const float rad = 0.1f, margin = 0.01f;
        const float shift = (rad + margin) * 2.f, centerx = shift * num / 2.f, centery = shift / 2.f;
        for (int i = 0; i < num; ++i)
            for (int j = i; j < num; ++j) {
                const float fi = (float)i, fj = (float)(j - i);
                bodies.insert(RigidBodyDesc().translation({fi * shift / 2.f + fj * shift - centerx, fi * shift + centery, 0.f})
                                  .cuboid_collider({rad, rad, rad}, 1.f).build());
                half.push_back({rad, rad, rad});
                offset.push_back({0.f, 0.f, 0.f});
            }
        BoxPileContacts producer(bodies, half, offset, margin);
        size_t nm = 0, nc = 0, nm0 = 0, nc0 = 0;
        for (int k = 0; k < steps; ++k) {
            std::vector<ColliderContactManifold> manifolds = producer.generate(bodies, offset);
            nm = manifolds.size();
            nc = 0;
            for (const auto& m : manifolds) nc += m.contacts.size();
            if (k == 0) {
                nm0 = nm;
                nc0 = nc;
            }
            world.step(bodies, joints, manifolds);
        }
        nb2_stats st = world.solver.stats();
        float vmax = 0.f, top = 0.f;
        for (size_t i = 0; i < bodies.len(); ++i) {
            const Vector3 v = bodies.get(i)->linear_velocity();
            vmax = std::fmax(vmax, std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]));
            top = std::fmax(top, bodies.get(i)->position().translation[1]);
        }
        std::printf("pyramid3: bodies %zu initial manifolds %zu contacts %zu rows %zu | last step manifolds %zu contacts %zu rows %u | steps %d mode %s\n",
                    bodies.len(), nm0, nc0, 3 * nc0, nm, nc, st.n_rows_two_body + st.n_rows_ground, steps,
                    reference ? "reference" : "coloured");
        std::printf("  phases %u residual %.3e max_penetration %.4f kinetic_energy %.4e non_finite %u\n", st.n_phases_velocity,
                    st.residual_max, st.max_penetration, st.kinetic_energy, st.non_finite);
        std::printf("  max |v| %.4f top y %.4f solver %.3f ms (assembly %.3f velocity %.3f update %.3f position %.3f)\n", vmax, top,
                    world.counters.solver_time, world.counters.assembly_time, world.counters.velocity_resolution_time,
                    world.counters.velocity_update_time, world.counters.position_resolution_time);
        const bool ok = st.non_finite == 0 && nm0 == 1335 && nc0 == 5340 && vmax < 1.0f && top > 6.3f && top < 6.6f;
        std::printf("%s\n", ok ? "OK" : "FAILED");
        return ok ? 0 : 1;
    } catch (const SolverError& e) {
        std::fprintf(stderr, "solver error %d: %s\n", e.code, e.what());
        return 2;
    }
}
