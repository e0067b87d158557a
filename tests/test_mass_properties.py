"""The one closed form the reference itself pins: cuboid mass / angular inertia against the tetrahedral
integration of the same solid (reference test `test_inertia_tensor3`,
src/volumetric/volumetric_convex3.rs:292-346).  The integration is restated here in numpy from
volumetric_convex3.rs:13-143 (tetrahedron inertia about a point), :145-177 (volume and centre of mass)
and :179-205 (mass properties); `scenes.cuboid_mass_properties` -- what every scene builder of this repo
uses for its bodies -- follows volumetric_cuboid.rs:8-18, 47-77 and must agree with it on the
reference test's own constants (half extents 0.96, density 2.37689, eccentricity 10, epsilon 1e-8)."""
import numpy as np

from nphysics_b200 import scenes


def tetrahedron_unit_inertia_wrt_point(point, p1, p2, p3, p4):
    """volumetric_convex3.rs:13-143, off-diagonal placement included."""
    p = np.array([p1 - point, p2 - point, p3 - point, p4 - point], dtype=np.float64)
    x, y, z = p[:, 0], p[:, 1], p[:, 2]

    def diag(v):
        return sum(v[i] * v[j] for i in range(4) for j in range(i, 4))

    def prod(u, v):
        return sum(u[i] * v[j] * (2.0 if i == j else 1.0) for i in range(4) for j in range(4)) * 0.05

    dx, dy, dz = diag(x), diag(y), diag(z)
    a0, b0, c0 = (dy + dz) * 0.1, (dz + dx) * 0.1, (dx + dy) * 0.1
    a1, b1, c1 = prod(y, z), prod(x, z), prod(x, y)
    return np.array([[a0, -b1, -c1], [-b1, b0, -a1], [-c1, -a1, c0]])


def tetrahedron_volume(p1, p2, p3, p4):
    """ncollide utils::tetrahedron_volume: |det[p2-p1, p3-p1, p4-p1]| / 6."""
    return abs(np.linalg.det(np.array([p2 - p1, p3 - p1, p4 - p1]))) / 6.0


def convex_mesh_mass_properties(coords, tris, density):
    """volumetric_convex3.rs:145-205."""
    center = coords.mean(axis=0)
    vol, com = 0.0, np.zeros(3)
    for t in tris:
        p2, p3, p4 = coords[t[0]], coords[t[1]], coords[t[2]]
        v = tetrahedron_volume(center, p2, p3, p4)
        com += (center + p2 + p3 + p4) / 4.0 * v
        vol += v
    com /= vol
    itot = np.zeros((3, 3))
    for t in tris:
        p2, p3, p4 = coords[t[0]], coords[t[1]], coords[t[2]]
        itot += tetrahedron_unit_inertia_wrt_point(com, com, p2, p3, p4) * tetrahedron_volume(com, p2, p3, p4)
    return vol * density, com, itot * density


def cuboid_mesh(extents, offset):
    """ncollide procedural::cuboid: the 8 corners of a box of full `extents`, 12 triangles."""
    h = np.asarray(extents, dtype=np.float64) / 2.0
    coords = np.array([[sx * h[0], sy * h[1], sz * h[2]] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)]) + offset
    quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    tris = [(q[0], q[1], q[2]) for q in quads] + [(q[0], q[2], q[3]) for q in quads]
    return coords, tris


def test_inertia_tensor3_reference_constants():
    coords, tris = cuboid_mesh((2.0 - 0.08,) * 3, 10.0)
    m_mesh, com, i_mesh = convex_mesh_mass_properties(coords, tris, 2.37689)
    m_box, i_box = scenes.cuboid_mass_properties((0.96, 0.96, 0.96), 2.37689)
    assert np.allclose(com, [10.0, 10.0, 10.0], atol=1e-9)
    assert abs(m_mesh - m_box) <= 1e-8 * max(1.0, m_box)
    assert np.allclose(i_mesh, i_box, rtol=1e-8, atol=1e-8)


def test_cuboid_closed_form_of_the_benchmark_box():
    """box m = 8 rho r^3, I = m (4/12)(r^2 + r^2) per axis (volumetric_cuboid.rs:47-77)."""
    r, rho = 0.1, 1.0
    m, inertia = scenes.cuboid_mass_properties((r, r, r), rho)
    assert m == 8.0 * rho * r ** 3
    assert np.allclose(np.diag(inertia), m * (4.0 / 12.0) * (r * r + r * r), rtol=1e-15)
    assert np.count_nonzero(inertia - np.diag(np.diag(inertia))) == 0
    # an anisotropic box against the mesh integration
    coords, tris = cuboid_mesh((0.4, 2.4, 0.8), np.array([1.0, -2.0, 0.5]))
    m_mesh, _, i_mesh = convex_mesh_mass_properties(coords, tris, 0.3)
    m_box, i_box = scenes.cuboid_mass_properties((0.2, 1.2, 0.4), 0.3)
    assert abs(m_mesh - m_box) <= 1e-12
    assert np.allclose(i_mesh, i_box, rtol=1e-10, atol=1e-12)
