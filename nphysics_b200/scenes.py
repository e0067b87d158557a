"""Scene builders and the stand-in manifold producer for the BASELINE.json configs.

The reference takes its contact manifolds from ncollide (out of scope, SURVEY.md
section 2 row 6); here they come from an analytic face-face generator for
(near) axis-aligned cuboids: every touching face pair yields one manifold with
the 4 corners of the overlap rectangle as Plane/Point (or Point/Plane) contacts,
recomputed from the CURRENT poses each step with persistent feature pairs and
keys -- the same records ncollide's polyhedral clipping produces for these piles
(SURVEY.md appendix B, C).  Scene constants follow examples3d/{pyramid3,boxes3,
wall3,constraints3}.rs as cited per function.
"""
import numpy as np

from . import abi

DEFAULT_MARGIN = 0.01       # ColliderDesc::default_margin, src/object/collider.rs:457-479
LINEAR_PREDICTION = 0.001   # collider.rs:457-479
DEFAULT_FRICTION = 0.5      # BasicMaterial::default, src/material/basic_material.rs:56-60
DEFAULT_RESTITUTION = 0.0


def quat_rotate(q, v):
    """Rotate vectors v[...,3] by unit quaternions q[...,4] (i,j,k,w)."""
    qv = q[..., :3]
    w = q[..., 3:4]
    t = 2.0 * np.cross(qv, v)
    return v + w * t + np.cross(qv, t)


def cuboid_mass_properties(half_extents, density):
    """volumetric_cuboid.rs:10-77: m = rho*8*hx*hy*hz, I_x = m*(4hy^2+4hz^2)/12, ..."""
    hx, hy, hz = [float(x) for x in half_extents]
    m = density * 8.0 * hx * hy * hz
    ix = m * (4 * hy * hy + 4 * hz * hz) / 12.0
    iy = m * (4 * hx * hx + 4 * hz * hz) / 12.0
    iz = m * (4 * hx * hx + 4 * hy * hy) / 12.0
    return m, np.diag([ix, iy, iz])


class Scene:
    """bodies + per-body cuboid collider (half extents, collider offset) + joints."""

    def __init__(self, bodies, half_extents, coll_offset, joints=None, params=None, name="scene"):
        self.bodies = bodies
        self.half_extents = np.asarray(half_extents, dtype=np.float64)  # (n,3); 0 = no collider
        self.coll_offset = np.asarray(coll_offset, dtype=np.float64)    # (n,3) collider translation wrt body
        self.joints = joints if joints is not None else np.zeros(0, dtype=abi.joint_dtype)
        self.params = params if params is not None else abi.default_params()
        self.name = name
        self.margin = DEFAULT_MARGIN
        self.friction = DEFAULT_FRICTION
        self.restitution = DEFAULT_RESTITUTION

    @property
    def n_dynamic(self):
        return int((self.bodies["status"] == abi.BODY_DYNAMIC).sum())


def _make_boxes(centers, rad, density, ground_half, ground_offset=(0.0, -0.2, 0.0)):
    """Body 0 = Ground (ground.rs) with a cuboid collider; bodies 1.. = dynamic cubes."""
    centers = np.asarray(centers, dtype=np.float64)
    n = len(centers) + 1
    bodies = abi.new_bodies(n)
    bodies["status"][0] = abi.BODY_STATIC
    bodies["flags"][0] = 0
    m, inertia = cuboid_mass_properties((rad, rad, rad), density)
    bodies["position"][1:, :3] = centers
    bodies["mass"][1:] = m
    bodies["local_inertia"][1:] = inertia.reshape(9)
    he = np.full((n, 3), rad)
    he[0] = ground_half
    off = np.zeros((n, 3))
    off[0] = ground_offset
    return bodies, he, off


def pyramid3(num=30, rad=0.1, density=1.0):
    """examples3d/pyramid3.rs:21-72: rows of num, num-1, ..., 1 cubes."""
    shift = (rad + DEFAULT_MARGIN) * 2.0
    centerx = shift * num / 2.0
    centery = shift / 2.0
    cs = []
    for i in range(num):
        for j in range(i, num):
            fi, fj = float(i), float(j - i)
            cs.append((fi * shift / 2.0 + fj * shift - centerx, fi * shift + centery, 0.0))
    bodies, he, off = _make_boxes(cs, rad, density, (6.0, 0.2, 6.0))
    return Scene(bodies, he, off, name="pyramid3")


def boxes3(nx=6, ny=6, nz=6, rad=0.1, density=1.0, height=0.0, ground_half=None, jitter=0.0, seed=12345):
    """examples3d/boxes3.rs:48-60 generalised to an nx*ny*nz grid (BASELINE config 2:
    50x40x50 settled, height=0)."""
    shift = (rad + DEFAULT_MARGIN) * 2.0
    i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    x = i * shift - shift * nx / 2.0
    y = j * shift + shift / 2.0 + height
    z = k * shift - shift * nz / 2.0
    cs = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1)
    if jitter > 0:
        rng = np.random.default_rng(seed)
        cs = cs + rng.uniform(-jitter, jitter, size=cs.shape)
    if ground_half is None:
        ground_half = (max(3.0, shift * nx / 2.0 + 2.5), 0.2, max(3.0, shift * nz / 2.0 + 2.5))
    bodies, he, off = _make_boxes(cs, rad, density, ground_half)
    return Scene(bodies, he, off, name="boxes3_%dx%dx%d" % (nx, ny, nz))


def wall3(width=50, height=10, rad=0.1, density=1.0):
    """examples3d/wall3.rs:43-69."""
    shift = (rad + DEFAULT_MARGIN) * 2.0
    centerx = shift * width / 2.0
    centery = shift / 2.0
    cs = [(i * shift - centerx, j * shift + centery, 0.0) for i in range(width) for j in range(height)]
    bodies, he, off = _make_boxes(cs, rad, density, (max(15.0, centerx + 2.0), 0.2, 3.0))
    return Scene(bodies, he, off, name="wall3_%dx%d" % (width, height))


def tile(scene, copies, pitch=20.0, per_row=64, first_world=0):
    """BASELINE config 5: `copies` independent translated copies of one scene (each with its own
    ground body), world w shifted by (w % per_row * pitch, 0, w // per_row * pitch); `first_world`
    selects worlds first_world .. first_world + copies - 1 of the lattice (e.g. its far corner)."""
    n = len(scene.bodies)
    bodies = np.tile(scene.bodies, copies)
    w = np.repeat(np.arange(first_world, first_world + copies), n)
    bodies["position"][:, 0] += (w % per_row) * pitch
    bodies["position"][:, 2] += (w // per_row) * pitch
    he = np.tile(scene.half_extents, (copies, 1))
    off = np.tile(scene.coll_offset, (copies, 1))
    joints = np.tile(scene.joints, copies)
    if len(joints):
        jw = np.repeat(np.arange(copies), len(scene.joints))
        joints["body1"] += (jw * n).astype(np.int32)
        joints["body2"] += (jw * n).astype(np.int32)
    s = Scene(bodies, he, off, joints, scene.params.copy(), name="%s_x%d" % (scene.name, copies))
    s.tile_of = (scene, copies)
    return s


def joint_chains(n_chains=10, links=6, rad=0.2, density=1.0, kind="mixed", ground_y=-5.2, pitch=3.0,
                 with_ground_collider=True):
    """BASELINE config 4: chains of `links` cubes hanging from the ground body.  Even chains use
    RevoluteConstraint links (examples3d/constraints3.rs:61-95: anchor offset (0,0,0.8), axis x),
    odd chains BallConstraint links (:138-169: offsets (cos a, 0.3, sin a) * 1.0).  Chains sit on a
    square lattice with `pitch` metres between them; the ground collider lies `ground_y` below."""
    per_row = int(np.ceil(np.sqrt(n_chains)))
    n = 1 + n_chains * links
    bodies = abi.new_bodies(n)
    bodies["status"][0] = abi.BODY_STATIC
    bodies["flags"][0] = 0
    m, inertia = cuboid_mass_properties((rad, rad, rad), density)
    bodies["mass"][1:] = m
    bodies["local_inertia"][1:] = inertia.reshape(9)
    he = np.full((n, 3), rad)
    span = per_row * pitch / 2.0 + 5.0
    he[0] = (span, 0.2, span) if with_ground_collider else (0.0, 0.0, 0.0)
    off = np.zeros((n, 3))
    off[0] = (0.0, ground_y - 0.2, 0.0)
    joints = abi.new_joints(n_chains * links)
    jn = 0
    for c in range(n_chains):
        ox = (c % per_row) * pitch - per_row * pitch / 2.0
        oz = (c // per_row) * pitch - per_row * pitch / 2.0
        revolute = (kind == "revolute") or (kind == "mixed" and c % 2 == 0)
        parent = 0
        parent_pos = np.array([ox, 0.0, oz])
        for l in range(links):
            b = 1 + c * links + l
            a1 = parent_pos if parent == 0 else np.zeros(3)
            if revolute:
                # constraints3.rs:66-95: body_anchor = z * (rad*3 + 0.2); pos -= body_anchor
                a2 = np.array([0.0, 0.0, 1.0]) * (rad * 3.0 + 0.2)
                joints["type"][jn] = abi.JOINT_REVOLUTE
                joints["axis1"][jn] = (1.0, 0.0, 0.0)
                joints["axis2"][jn] = (1.0, 0.0, 0.0)
            else:
                # constraints3.rs:138-169: first link sits on the anchor, the others hang off
                # (cos a, 0.3, sin a) * rad * 5 with a = i * 2 pi / num
                ang = l * 2.0 * np.pi / links
                a2 = np.zeros(3) if l == 0 else np.array([np.cos(ang), 0.3, np.sin(ang)]) * (rad * 5.0)
                joints["type"][jn] = abi.JOINT_BALL
            pos = parent_pos - a2
            bodies["position"][b, :3] = pos
            joints["body1"][jn] = parent
            joints["body2"][jn] = b
            joints["anchor1"][jn] = a1
            joints["anchor2"][jn] = a2
            jn += 1
            parent = b
            parent_pos = pos
    return Scene(bodies, he, off, joints, name="chains_%dx%d_%s" % (n_chains, links, kind))


def ragdolls(n=10, pitch=3.0, height=5.0, spin=2.0, density=0.3):
    """BASELINE config 4, second half: the topology of examples3d/ragdoll3.rs:65-135 (torso + head + two
    arms + two legs joined by five spherical joints) restated with BallConstraints between rigid bodies
    (the shipped example builds a Multibody, which is out of scope).  Members are cuboids of the
    members' extents (torso 0.2 x 1.2 x 0.4, head 0.4, arms 0.15 x 1.05, legs 0.15 x 1.55), density 0.3,
    `space` 0.15 between members.  Ragdolls sit on a square lattice `height` above the origin plane and
    start with an angular velocity of `spin` rad/s about z on the torso, so that the joints work while
    the whole figure falls.  No colliders (joint rows only)."""
    body_rady, body_radz, body_radx = 0.6, 0.2, 0.1
    head_rad, member_rad, arm_length, leg_length, space = 0.2, 0.075, 0.45, 0.7, 0.15
    members = [  # half extents, anchor on the torso (parent_shift), anchor on the member (body_shift)
        ((head_rad, head_rad, head_rad), (0.0, body_rady + head_rad + space, 0.0), (0.0, 0.0, 0.0)),
        ((member_rad, arm_length + member_rad, member_rad), (0.0, body_rady, body_radx + 2.0 * space), (0.0, arm_length + space, 0.0)),
        ((member_rad, arm_length + member_rad, member_rad), (0.0, body_rady, -body_radx - 2.0 * space), (0.0, arm_length + space, 0.0)),
        ((member_rad, leg_length + member_rad, member_rad), (0.0, -body_rady, body_radx), (0.0, leg_length + space, 0.0)),
        ((member_rad, leg_length + member_rad, member_rad), (0.0, -body_rady, -body_radx), (0.0, leg_length + space, 0.0)),
    ]
    per_row = int(np.ceil(np.sqrt(n)))
    nb = 1 + 6 * n
    bodies = abi.new_bodies(nb)
    bodies["status"][0] = abi.BODY_STATIC
    bodies["flags"][0] = 0
    he = np.zeros((nb, 3))
    off = np.zeros((nb, 3))
    joints = abi.new_joints(5 * n, abi.JOINT_BALL)
    for r in range(n):
        origin = np.array([(r % per_row) * pitch, height, (r // per_row) * pitch])
        torso = 1 + 6 * r
        m, inertia = cuboid_mass_properties((body_radx, body_rady, body_radz), density)
        bodies["position"][torso, :3] = origin
        bodies["mass"][torso] = m
        bodies["local_inertia"][torso] = inertia.reshape(9)
        bodies["velocity"][torso, 5] = spin
        for k, (half, a1, a2) in enumerate(members):
            b = torso + 1 + k
            m, inertia = cuboid_mass_properties(half, density)
            bodies["position"][b, :3] = origin + np.array(a1) - np.array(a2)
            bodies["mass"][b] = m
            bodies["local_inertia"][b] = inertia.reshape(9)
            j = 5 * r + k
            joints["body1"][j] = torso
            joints["body2"][j] = b
            joints["anchor1"][j] = a1
            joints["anchor2"][j] = a2
    return Scene(bodies, he, off, joints, name="ragdolls_%d" % n)


# ----------------------------------------------------------------------------------------------
# the manifold producer
# ----------------------------------------------------------------------------------------------
class ContactGenerator:
    """Persistent face-face feature pairs between (near) axis-aligned cuboids.

    Feature pairs are discovered once from the initial poses (touching or within the prediction
    distance); `generate(positions)` then re-evaluates every contact of every pair at the given
    poses, drops contacts farther apart than the narrow-phase prediction
    (margin+linear_prediction on both sides = 0.022, collider.rs:542-549) and emits
    (nb2_manifold[], nb2_contact[]) with stable keys.
    """

    def __init__(self, scene, flip_fraction=0.0, search=None, order="pair"):
        """order="pair": pairs sorted by (a, b).  order="owner": the canonical order of the device producer
        (csrc/narrowphase.cu) -- every pair is owned by its dynamic collider (the lower index of two dynamic
        ones) and pairs are listed owner by owner, an owner's non-dynamic partners before its dynamic ones."""
        self.scene = scene
        pos = scene.bodies["position"].astype(np.float64)
        centers = pos[:, :3] + scene.coll_offset
        he = scene.half_extents
        status = scene.bodies["status"]
        has_coll = he.max(axis=1) > 0
        margin = scene.margin
        reach = 2.0 * (margin + LINEAR_PREDICTION)
        movable = (status == abi.BODY_DYNAMIC) | (status == abi.BODY_MULTIBODY_LINK)  # a multibody link collides like a body
        dyn = np.nonzero(movable & has_coll)[0]
        big = np.nonzero(~movable & has_coll)[0]
        pa, pb = [], []
        if len(dyn) > 1:
            from scipy.spatial import cKDTree
            r_max = float(he[dyn].max())
            radius = search if search is not None else (2.0 * r_max + reach) * np.sqrt(2.0) * 1.01
            tree = cKDTree(centers[dyn])
            pr = tree.query_pairs(radius, output_type="ndarray")
            pa.append(dyn[pr[:, 0]])
            pb.append(dyn[pr[:, 1]])
        for g in big:  # ground-like bodies against every dynamic box
            pa.append(np.full(len(dyn), g))
            pb.append(dyn)
        if pa:
            a = np.concatenate(pa)
            b = np.concatenate(pb)
        else:
            a = np.zeros(0, dtype=np.int64)
            b = np.zeros(0, dtype=np.int64)
        d = centers[b] - centers[a]
        hsum = he[a] + he[b]
        gap = np.abs(d) - hsum                      # per axis; > 0 separated
        overlap = -gap                              # per axis overlap extent
        eps = 1e-6
        keep = np.zeros(len(a), dtype=bool)
        axis = np.zeros(len(a), dtype=np.int64)
        for ax in range(3):
            o1, o2 = (ax + 1) % 3, (ax + 2) % 3
            ok = (gap[:, ax] <= reach + eps) & (gap[:, ax] > -0.5 * np.minimum(he[a, ax], he[b, ax])) & \
                 (overlap[:, o1] > eps) & (overlap[:, o2] > eps)
            ok &= ~keep
            axis[ok] = ax
            keep |= ok
        a, b, d, axis = a[keep], b[keep], d[keep], axis[keep]
        # deterministic order: by (a, b), or owner by owner
        if order == "owner":
            a_dyn = status[a] == abi.BODY_DYNAMIC
            owner = np.where(a_dyn, a, b)
            partner = np.where(a_dyn, b, a)
            perm = np.lexsort((partner, a_dyn, owner))
        else:
            perm = np.lexsort((b, a))
        a, b, d, axis = a[perm], b[perm], d[perm], axis[perm]
        npairs = len(a)
        rows = np.arange(npairs)
        sign = np.where(d[rows, axis] >= 0, 1.0, -1.0)
        ca, cb = centers[a], centers[b]
        u = (axis + 1) % 3
        v = (axis + 2) % 3
        lo_u = np.maximum(ca[rows, u] - he[a, u], cb[rows, u] - he[b, u])
        hi_u = np.minimum(ca[rows, u] + he[a, u], cb[rows, u] + he[b, u])
        lo_v = np.maximum(ca[rows, v] - he[a, v], cb[rows, v] - he[b, v])
        hi_v = np.minimum(ca[rows, v] + he[a, v], cb[rows, v] + he[b, v])
        corners_u = np.stack([lo_u, hi_u, hi_u, lo_u], axis=1)   # (npairs,4)
        corners_v = np.stack([lo_v, lo_v, hi_v, hi_v], axis=1)
        # local points in the COLLIDER frames of a (on its face) and b (on its face)
        la = np.zeros((npairs, 4, 3))
        lb = np.zeros((npairs, 4, 3))
        for k in range(4):
            la[rows, k, u] = corners_u[:, k] - ca[rows, u]
            la[rows, k, v] = corners_v[:, k] - ca[rows, v]
            la[rows, k, axis] = sign * he[a, axis]
            lb[rows, k, u] = corners_u[:, k] - cb[rows, u]
            lb[rows, k, v] = corners_v[:, k] - cb[rows, v]
            lb[rows, k, axis] = -sign * he[b, axis]
        na = np.zeros((npairs, 3))
        na[rows, axis] = sign                       # outward normal of a's face, local to a
        # flipped pairs: body1 = b carries the Point, body2 = a carries the Plane
        h = (a * 2654435761 + b * 40503) % 1000
        self.flip = h < int(flip_fraction * 1000)
        self.a, self.b = a, b
        self.la, self.lb, self.na = la, lb, na
        self.npairs = npairs
        self.reach = reach

    def generate(self, positions=None):
        sc = self.scene
        if positions is None:
            positions = sc.bodies["position"]
        positions = np.asarray(positions, dtype=np.float64)
        a, b = self.a, self.b
        npairs = self.npairs
        off = sc.coll_offset
        qa, qb = positions[a, 3:7], positions[b, 3:7]
        # collider pose = body pose * translation(offset)
        ta = positions[a, :3] + quat_rotate(qa, off[a])
        tb = positions[b, :3] + quat_rotate(qb, off[b])
        wa = ta[:, None, :] + quat_rotate(qa[:, None, :], self.la)     # points on a's face
        wb = tb[:, None, :] + quat_rotate(qb[:, None, :], self.lb)     # points on b's face
        n = quat_rotate(qa, self.na)                                   # world normal a -> b
        depth = -np.einsum("pj,pkj->pk", n, wb - wa)                   # Plane(a)/Point(b)
        w1 = wb + n[:, None, :] * depth[:, :, None]                    # projection on a's plane
        keep = depth > -self.reach
        nper = keep.sum(axis=1)
        mkeep = nper > 0
        flip = self.flip
        m_idx = np.nonzero(mkeep)[0]
        nm = len(m_idx)
        manifolds = np.zeros(nm, dtype=abi.manifold_dtype)
        body1 = np.where(flip, b, a)
        body2 = np.where(flip, a, b)
        manifolds["body1"] = body1[m_idx]
        manifolds["body2"] = body2[m_idx]
        manifolds["num_contacts"] = nper[m_idx]
        first = np.concatenate([[0], np.cumsum(nper[m_idx])[:-1]]) if nm else np.zeros(0)
        manifolds["first_contact"] = first
        manifolds["margin1"] = sc.margin
        manifolds["margin2"] = sc.margin
        manifolds["friction"] = sc.friction
        manifolds["restitution"] = sc.restitution
        manifolds["coll1_wrt_body"][:, 6] = 1.0
        manifolds["coll2_wrt_body"][:, 6] = 1.0
        manifolds["coll1_wrt_body"][:, :3] = off[body1[m_idx]]
        manifolds["coll2_wrt_body"][:, :3] = off[body2[m_idx]]
        pi, ki = np.nonzero(keep)          # row-major: pair order, corner order
        nc = len(pi)
        contacts = np.zeros(nc, dtype=abi.contact_dtype)
        f = flip[pi]
        nn = n[pi]
        contacts["normal"] = np.where(f[:, None], -nn, nn)
        contacts["depth"] = depth[pi, ki]
        contacts["world1"] = np.where(f[:, None], wb[pi, ki], w1[pi, ki])
        contacts["world2"] = np.where(f[:, None], w1[pi, ki], wb[pi, ki])
        contacts["key"] = (pi.astype(np.uint64) * np.uint64(4) + ki.astype(np.uint64) + np.uint64(1))
        contacts["local1"] = np.where(f[:, None], self.lb[pi, ki], self.la[pi, ki])
        contacts["local2"] = np.where(f[:, None], self.la[pi, ki], self.lb[pi, ki])
        nl = self.na[pi]
        contacts["dir1"] = np.where(f[:, None], 0.0, nl)
        contacts["dir2"] = np.where(f[:, None], nl, 0.0)
        contacts["geom1"] = np.where(f, abi.GEOM_POINT, abi.GEOM_PLANE)
        contacts["geom2"] = np.where(f, abi.GEOM_PLANE, abi.GEOM_POINT)
        return manifolds, contacts


def scene_colliders(scene):
    """nb2_collider records of a Scene: one cuboid per body that has one (in body order, so that collider
    indices order like body indices), translation-only collider offsets, the scene-wide margin and material."""
    has = scene.half_extents.max(axis=1) > 0
    idx = np.nonzero(has)[0]
    c = abi.new_colliders(len(idx))
    c["half_extents"] = scene.half_extents[idx]
    c["translation_wrt_body"] = scene.coll_offset[idx]
    c["margin"] = scene.margin
    c["friction"] = scene.friction
    c["restitution"] = scene.restitution
    c["body"] = idx
    return c


def row_counts(scene, manifolds):
    """(N_R2, N_RG): two-dynamic-body / ground velocity rows of the contact set (+ joints, using
    the rows each joint type emits when no limit is active)."""
    st = scene.bodies["status"]
    d1 = st[manifolds["body1"]] == abi.BODY_DYNAMIC
    d2 = st[manifolds["body2"]] == abi.BODY_DYNAMIC
    rows = 3 * manifolds["num_contacts"].astype(np.int64)
    r2 = int(rows[d1 & d2].sum())
    rg = int(rows[d1 ^ d2].sum())
    per_type = np.array([3, 5, 5, 4, 3, 4, 4, 4, 6, 3])
    j = scene.joints
    if len(j):
        jd1 = st[j["body1"]] == abi.BODY_DYNAMIC
        jd2 = st[j["body2"]] == abi.BODY_DYNAMIC
        jr = per_type[j["type"]]
        r2 += int(jr[jd1 & jd2].sum())
        rg += int(jr[jd1 ^ jd2].sum())
    return r2, rg


def joint_zoo(seed=7, rad=0.2, density=1.0, with_limits=True):
    """One short chain (ground -> b1 -> b2) per constraint-based joint type (the mix of
    examples3d/constraints3.rs, plus Fixed / Cylindrical / Cartesian for completeness), slightly
    perturbed so that every velocity row and every position generator has something to correct."""
    rng = np.random.default_rng(seed)
    types = list(range(10))
    n = 1 + 2 * len(types)
    bodies = abi.new_bodies(n)
    bodies["status"][0] = abi.BODY_STATIC
    bodies["flags"][0] = 0
    m, inertia = cuboid_mass_properties((rad, rad * 0.8, rad * 1.2), density)
    bodies["mass"][1:] = m
    bodies["local_inertia"][1:] = inertia.reshape(9)
    he = np.zeros((n, 3))
    off = np.zeros((n, 3))
    joints = abi.new_joints(2 * len(types))

    def unit(v):
        v = np.asarray(v, dtype=np.float64)
        return v / np.linalg.norm(v)

    def rand_quat(scale):
        w = rng.normal(size=3) * scale
        ang = np.linalg.norm(w)
        if ang < 1e-12:
            return np.array([0.0, 0.0, 0.0, 1.0])
        ax = w / ang
        return np.concatenate([ax * np.sin(ang / 2), [np.cos(ang / 2)]])

    jn = 0
    for k, t in enumerate(types):
        base = np.array([3.0 * k, 5.0, 0.0])
        parent, parent_pos = 0, base
        for l in range(2):
            b = 1 + 2 * k + l
            a2 = np.array([0.0, 0.0, 1.0]) * (rad * 3.0 + 0.2)
            pos = parent_pos - a2 + rng.normal(size=3) * 0.01     # small anchor error
            bodies["position"][b, :3] = pos
            bodies["position"][b, 3:] = rand_quat(0.05)           # small axis error
            bodies["velocity"][b] = rng.normal(size=6) * 0.3
            joints["type"][jn] = t
            joints["body1"][jn] = parent
            joints["body2"][jn] = b
            joints["anchor1"][jn] = parent_pos if parent == 0 else np.zeros(3)
            joints["anchor2"][jn] = a2
            ax = unit(rng.normal(size=3))
            joints["axis1"][jn] = ax
            joints["axis2"][jn] = ax
            if t == abi.JOINT_PIN_SLOT:
                joints["axis3"][jn] = unit(np.cross(ax, rng.normal(size=3)))
                joints["axis2"][jn] = joints["axis3"][jn]
            if t == abi.JOINT_UNIVERSAL:
                ax2 = unit(np.cross(ax, rng.normal(size=3)))
                joints["axis2"][jn] = ax2
                joints["angle"][jn] = np.pi / 2.0
            if t == abi.JOINT_PRISMATIC and with_limits:
                joints["flags"][jn] = abi.JOINT_FLAG_MIN_OFFSET | (abi.JOINT_FLAG_MAX_OFFSET if l else 0)
                joints["min_offset"][jn] = 0.05 if l == 0 else -0.4
                joints["max_offset"][jn] = 0.3
            if t in (abi.JOINT_FIXED, abi.JOINT_CARTESIAN):
                joints["ref_frame1"][jn] = rand_quat(0.3)
                joints["ref_frame2"][jn] = joints["ref_frame1"][jn]
            jn += 1
            parent, parent_pos = b, pos
    return Scene(bodies, he, off, joints, name="joint_zoo")


# ----------------------------------------------------------------------------------------------
# reduced-coordinate multibodies
# ----------------------------------------------------------------------------------------------
class MultibodyBuilder:
    """MultibodyDesc restated for the flat ABI (src/object/multibody.rs:1311-1470): `add` appends a link to the
    current multibody (its NB2_BODY_MULTIBODY_LINK record is appended to the body set), `finish` closes it."""

    def __init__(self, bodies, half_extents, coll_offset):
        self.bodies = list(bodies)
        self.he = [np.asarray(h, dtype=np.float64) for h in half_extents]
        self.off = [np.asarray(o, dtype=np.float64) for o in coll_offset]
        self.links = []
        self.multibodies = []
        self._first = 0

    def add(self, parent, joint_type, half_extents, density, parent_shift=(0, 0, 0), body_shift=(0, 0, 0), axis=(1, 0, 0),
            coords=None, velocity=None, collider=True, **kw):
        b = abi.new_bodies(1)[0].copy()
        b["status"] = abi.BODY_MULTIBODY_LINK
        m, inertia = cuboid_mass_properties(half_extents, density)
        b["mass"] = m
        b["local_inertia"] = inertia.reshape(9)
        self.bodies.append(b)
        self.he.append(np.asarray(half_extents, dtype=np.float64) if collider else np.zeros(3))
        self.off.append(np.zeros(3))
        l = abi.new_mb_links(1, joint_type)
        l["multibody"] = len(self.multibodies)
        l["parent"] = parent
        l["body"] = len(self.bodies) - 1
        l["parent_shift"] = parent_shift
        l["body_shift"] = body_shift
        a = np.asarray(axis, dtype=np.float64)
        l["axis"] = a / np.linalg.norm(a)
        if coords is not None:
            l["coords"][0, :len(coords)] = coords
        if velocity is not None:
            l["velocity"][0, :len(velocity)] = velocity
        for k, v in kw.items():
            l[k] = v
        self.links.append(l[0].copy())
        return len(self.links) - 1 - self._first

    def finish(self, gravity=True):
        mb = np.zeros(1, dtype=abi.multibody_dtype)
        mb["first_link"] = self._first
        mb["n_links"] = len(self.links) - self._first
        mb["flags"] = abi.BODY_FLAG_GRAVITY if gravity else 0
        self.multibodies.append(mb[0].copy())
        self._first = len(self.links)

    def _forward_kinematics(self, bodies):
        """Link poses from the joint coordinates (Multibody::update_kinematics, multibody.rs:830-863) so that the link
        records start where the library will put them (the contact producers look at them before the first step)."""
        def qmul(a, b):
            ai, aj, ak, aw = a
            bi, bj, bk, bw = b
            return np.array([aw * bi + ai * bw + aj * bk - ak * bj, aw * bj - ai * bk + aj * bw + ak * bi,
                             aw * bk + ai * bj - aj * bi + ak * bw, aw * bw - ai * bi - aj * bj - ak * bk])
        world = {}
        for mb in self.multibodies:
            for k in range(int(mb["n_links"])):
                l = self.links[int(mb["first_link"]) + k]
                ps, bs = l["parent_shift"].astype(np.float64), l["body_shift"].astype(np.float64)
                c = l["coords"].astype(np.float64)
                jt = int(l["joint_type"])
                if jt == abi.MBJ_FREE:
                    t, q = c[:3], c[3:7]
                elif jt == abi.MBJ_FIXED:
                    q = c[3:7]
                    t = ps + c[:3] + quat_rotate(q, bs)
                elif jt == abi.MBJ_BALL:
                    q = c[:4]
                    t = ps - quat_rotate(q, bs)
                elif jt == abi.MBJ_REVOLUTE:
                    ax = l["axis"].astype(np.float64)
                    q = np.append(ax * np.sin(c[0] / 2), np.cos(c[0] / 2))
                    t = ps - quat_rotate(q, bs)
                else:
                    q = np.array([0.0, 0.0, 0.0, 1.0])
                    t = ps - bs + l["axis"].astype(np.float64) * c[0]
                if k > 0:
                    pt, pq = world[(int(mb["first_link"]), int(l["parent"]))]
                    t, q = pt + quat_rotate(pq, t), qmul(pq, q)
                world[(int(mb["first_link"]), k)] = (t, q)
                bodies["position"][int(l["body"]), :3] = t
                bodies["position"][int(l["body"]), 3:] = q

    def scene(self, name, joints=None):
        bodies = np.array(self.bodies, dtype=abi.body_dtype)
        self._forward_kinematics(bodies)
        sc = Scene(bodies, np.array(self.he), np.array(self.off), joints, name=name)
        sc.multibodies = np.array(self.multibodies, dtype=abi.multibody_dtype)
        sc.mb_links = np.array(self.links, dtype=abi.mb_link_dtype)
        return sc


def _ground_only(ground_half=(20.0, 0.2, 20.0)):
    bodies = abi.new_bodies(1)
    bodies["status"][0] = abi.BODY_STATIC
    bodies["flags"][0] = 0
    return MultibodyBuilder(bodies, [ground_half], [(0.0, -0.2, 0.0)])


def multibody_ragdolls(n=10, pitch=3.0, height=5.0, spin=2.0, density=0.3, colliders=True):
    """examples3d/ragdoll3.rs:65-160 AS SHIPPED: one Multibody per ragdoll -- a FreeJoint torso and five BallJoint
    members (head, two arms, two legs) with the example's parent / body shifts: 6 links, 21 dofs.  Members are
    cuboids of the members' extents (the example's balls and capsules are ncollide shapes).  `spin`: initial angular
    velocity of the torso about z."""
    body_rady, body_radz, body_radx = 0.6, 0.2, 0.1
    head_rad, member_rad, arm_length, leg_length, space = 0.2, 0.075, 0.45, 0.7, 0.15
    members = [
        ((head_rad, head_rad, head_rad), (0.0, body_rady + head_rad + space, 0.0), (0.0, 0.0, 0.0)),
        ((member_rad, arm_length + member_rad, member_rad), (0.0, body_rady, body_radx + 2.0 * space), (0.0, arm_length + space, 0.0)),
        ((member_rad, arm_length + member_rad, member_rad), (0.0, body_rady, -body_radx - 2.0 * space), (0.0, arm_length + space, 0.0)),
        ((member_rad, leg_length + member_rad, member_rad), (0.0, -body_rady, body_radx), (0.0, leg_length + space, 0.0)),
        ((member_rad, leg_length + member_rad, member_rad), (0.0, -body_rady, -body_radx), (0.0, leg_length + space, 0.0)),
    ]
    per_row = int(np.ceil(np.sqrt(n)))
    mb = _ground_only((max(20.0, (per_row + 1) * pitch), 0.2, max(20.0, (per_row + 1) * pitch)))
    for r in range(n):
        origin = [(r % per_row) * pitch, height, (r // per_row) * pitch]
        root = mb.add(-1, abi.MBJ_FREE, (body_radx, body_rady, body_radz), density, coords=origin + [0, 0, 0, 1],
                      velocity=[0, 0, 0, 0, 0, spin], collider=colliders)
        for half, a1, a2 in members:
            mb.add(root, abi.MBJ_BALL, half, density, parent_shift=a1, body_shift=a2, collider=colliders)
        mb.finish()
    return mb.scene("multibody_ragdolls_%d" % n)


def multibody_chain(joint_type=None, links=4, rad=0.2, density=1.0, anchor=(0.0, 5.0, 0.0), spacing=0.8, axis=(1, 0, 0),
                    damping=None, root_fixed=True, **kw):
    """A chain hanging from the ground in the manner of examples3d/multibody3.rs: link k hangs `spacing` below link
    k-1 along -z (so that it swings under gravity about x); the root is attached to the world by the same joint."""
    joint_type = abi.MBJ_REVOLUTE if joint_type is None else joint_type
    mb = _ground_only()
    for k in range(links):
        extra = dict(kw)
        if damping is not None:
            extra["damping"] = [damping] * 6
        if k == 0:
            mb.add(-1, joint_type, (rad, rad, rad), density, parent_shift=anchor, body_shift=(0, 0, spacing), axis=axis, **extra)
        else:
            mb.add(k - 1, joint_type, (rad, rad, rad), density, parent_shift=(0, 0, 0), body_shift=(0, 0, spacing), axis=axis,
                   **extra)
    mb.finish()
    return mb.scene("multibody_chain")
