// examples3d/ragdoll3.rs over the C++ host mirror, collision detection included: the ground collider
// (ragdoll3.rs:31-42) and one collider per link go into a DefaultColliderSet and the loop is the reference's
// `mechanical_world.step(&mut geometrical_world, &mut bodies, &mut colliders, &mut joint_constraints, ..)`.
// The geometrical world here is the device manifold producer: cuboids only, so the head (a Ball in
// the reference) and the limbs (Capsules) collide as their bounding cuboids -- the same extents the mass
// properties of example_ragdoll3.cpp use.  Checked: the joints do not drift, nothing falls through the ground,
// every state stays finite.
//
// Build:  g++ -O2 -std=c++17 example_ragdoll3_colliders.cpp -o example_ragdoll3_colliders -L.. -lnphysics_b200 -Wl,-rpath,'$ORIGIN/..'
// Usage:  ./example_ragdoll3_colliders [steps] [n]      (n^3 ragdolls, default 2)
#include <cstdio>
#include <cstdlib>

#include "nphysics_b200.hpp"

using namespace nphysics;

static Vector3 rotate(const Quaternion& q, const Vector3& p) {
    const float vx = q[0], vy = q[1], vz = q[2], w = q[3];
    const float tx = 2.f * (vy * p[2] - vz * p[1]), ty = 2.f * (vz * p[0] - vx * p[2]), tz = 2.f * (vx * p[1] - vy * p[0]);
    return {p[0] + w * tx + (vy * tz - vz * ty), p[1] + w * ty + (vz * tx - vx * tz), p[2] + w * tz + (vx * ty - vy * tx)};
}

int main(int argc, char** argv) {
    const int steps = argc > 1 ? std::atoi(argv[1]) : 150;
    const int n = argc > 2 ? std::atoi(argv[2]) : 2;
    const float body_rady = 0.6f, body_radz = 0.2f, body_radx = 0.1f, head_rad = 0.2f, member_rad = 0.075f;
    const float arm_length = 0.45f, leg_length = 0.7f, space = 0.15f, density = 0.3f;
    struct Member { Vector3 half, parent_shift, body_shift; };
    const Member members[5] = {
        {{head_rad, head_rad, head_rad}, {0.f, body_rady + head_rad + space, 0.f}, {0.f, 0.f, 0.f}},
        {{member_rad, arm_length + member_rad, member_rad}, {0.f, body_rady, body_radx + 2.f * space}, {0.f, arm_length + space, 0.f}},
        {{member_rad, arm_length + member_rad, member_rad}, {0.f, body_rady, -body_radx - 2.f * space}, {0.f, arm_length + space, 0.f}},
        {{member_rad, leg_length + member_rad, member_rad}, {0.f, -body_rady, body_radx}, {0.f, leg_length + space, 0.f}},
        {{member_rad, leg_length + member_rad, member_rad}, {0.f, -body_rady, -body_radx}, {0.f, leg_length + space, 0.f}}};
    try {
        MechanicalWorld world({0.f, -9.81f, 0.f});
        DefaultBodySet bodies;
        DefaultJointConstraintSet joints;
        DefaultColliderSet colliders;
        GeometricalWorld geometrical_world;
        const float ground_thickness = 0.2f;
        const DefaultBodyHandle ground = bodies.insert(Ground::make());
        // The ragdolls arrive at 8 - 12 m/s, 0.13 - 0.2 m per step: the colliders' linear prediction (collider.rs:457-479,
        // 0.001 by default) is raised so that a contact exists the step BEFORE the faces meet (there is no CCD here).
        const float prediction = 0.15f;
        colliders.insert(ColliderDesc({5.f, ground_thickness, 5.f}).translation({0.f, -ground_thickness, 0.f}).linear_prediction(prediction).build({ground, 0}));
        std::vector<DefaultBodyHandle> handles;
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j)
                for (int k = 0; k < n; ++k) {
                    Isometry3 at;
                    at.translation = {i * 1.f - n * 0.5f, j * 5.f + 3.f, k * 1.f - n * 0.5f};
                    at.rotation = {0.f, 0.f, 0.f, 1.f};
                    FreeJoint free(at);
                    MultibodyDesc body(free);
                    body.cuboid({body_radx, body_rady, body_radz}, density);
                    for (const Member& m : members)
                        body.add_child(BallJoint()).set_parent_shift(m.parent_shift).set_body_shift(m.body_shift).cuboid(m.half, density);
                    const DefaultBodyHandle h = bodies.insert_multibody(Multibody(body));
                    handles.push_back(h);
                    colliders.insert(ColliderDesc({body_radx, body_rady, body_radz}).linear_prediction(prediction).build({h, 0}));
                    for (size_t m = 0; m < 5; ++m) colliders.insert(ColliderDesc(members[m].half).linear_prediction(prediction).build({h + 1 + m, 0}));
                }
        float worst = 0.f, lowest = 1e9f;
        for (int s = 0; s < steps; ++s) {
            world.step(geometrical_world, bodies, colliders, joints);
            for (size_t r = 0; r < bodies.num_multibodies(); ++r) {
                const Multibody& mb = bodies.multibody(r);
                const Isometry3 torso = mb.link_position(0);
                for (int k = 0; k < 5; ++k) {  // joint anchor seen from the torso and from the member
                    const Isometry3 member = mb.link_position(1 + k);
                    const Vector3 a = rotate(torso.rotation, members[k].parent_shift), b = rotate(member.rotation, members[k].body_shift);
                    for (int c = 0; c < 3; ++c) {
                        const float gap = std::fabs((torso.translation[c] + a[c]) - (member.translation[c] + b[c]));
                        worst = gap > worst ? gap : worst;
                    }
                }
                lowest = torso.translation[1] < lowest ? torso.translation[1] : lowest;
            }
        }
        nb2_stats st = world.solver.stats();
        std::printf("ragdoll3 with colliders: %zu multibodies (%zu links, %zu colliders, %u pairs now), %d steps, largest joint gap %.3e m, "
                    "lowest torso y %.3f, non-finite %u\n",
                    bodies.num_multibodies(), bodies.num_multibodies() * 6, colliders.len(), geometrical_world.n_pairs, steps, worst,
                    lowest, st.non_finite);
        return (worst < 1e-4f && st.non_finite == 0 && lowest > -0.5f) ? 0 : 1;
    } catch (const SolverError& e) {
        std::fprintf(stderr, "solver error %d: %s\n", e.code, e.what());
        return 2;
    }
}
