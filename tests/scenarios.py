"""Seeded synthetic rod configurations shared by the tests, smoke() and bench.py (harness code)."""
import numpy as np


def quat_from_z_to(direction):
    """Eigen::Quaterniond::FromTwoVectors((0,0,1), direction) for an array of directions -> (x,y,z,w).
    Used by the reference when reading rod files (SylinderSystem.cpp:337-339)."""
    d = np.asarray(direction, dtype=np.float64)
    d = d / np.linalg.norm(d, axis=-1, keepdims=True)
    c = d[..., 2]
    axis = np.stack([-d[..., 1], d[..., 0], np.zeros_like(c)], axis=-1)  # z x d
    s = np.sqrt((1.0 + c) * 2.0)
    q = np.zeros(d.shape[:-1] + (4,))
    ok = c > -1.0 + 1e-12
    q[ok, :3] = axis[ok] / s[ok, None]
    q[ok, 3] = s[ok] * 0.5
    q[~ok] = np.array([1.0, 0.0, 0.0, 0.0])  # antiparallel: rotate pi about x
    return q


def random_quat(rng, n):
    q = rng.normal(size=(n, 4))
    return q / np.linalg.norm(q, axis=1)[:, None]


def random_rods(n, box, length=0.25, radius=0.0125, seed=0, lo=0.0, frac_sphere=0.0, frac_immovable=0.0,
                aligned=None, length_sigma=0.0):
    """Uniform positions in [lo, lo+box]^3 (box may be a 3-vector), isotropic (or aligned) orientations."""
    rng = np.random.default_rng(seed)
    box = np.broadcast_to(np.asarray(box, dtype=np.float64), (3,))
    pos = lo + rng.uniform(0, 1, size=(n, 3)) * box
    if aligned is None:
        quat = random_quat(rng, n)
    else:
        d = np.asarray(aligned, dtype=np.float64) + aligned_noise(rng, n, 0.1)
        quat = quat_from_z_to(d)
    gid = rng.permutation(n).astype(np.int32)
    L = np.full(n, float(length))
    if length_sigma > 0:
        L = L * np.exp(rng.normal(0, length_sigma, size=n))
    R = np.full(n, float(radius))
    if frac_sphere > 0:
        sp = rng.uniform(size=n) < frac_sphere
        L[sp] = R[sp] * rng.uniform(0.0, 1.9, size=sp.sum())  # length < 2 radius -> treated as sphere
        R[sp] *= 3.0
    imm = (rng.uniform(size=n) < frac_immovable).astype(np.uint8)
    return dict(gid=gid, pos=pos, quat=quat, length=L, radius=R, immovable=imm)


def aligned_noise(rng, n, sigma):
    return rng.normal(0, sigma, size=(n, 3))


def box_for_volume_fraction(n, length, radius, phi):
    vol = np.pi * radius**2 * length + 4.0 / 3.0 * np.pi * radius**3
    return float((n * vol / phi) ** (1.0 / 3.0))


def thermal_velocity(rods, viscosity, dt, kbt=0.00411, seed=1):
    """Synthetic non-constraint velocity: Brownian-scale kicks sqrt(2 kBT/(zeta dt)) N(0,1) per rod dof
    (the role velocityBrown plays in SylinderSystem::calcVelocityNonCon, SylinderSystem.cpp:724-800)."""
    rng = np.random.default_rng(seed)
    n = len(rods["gid"])
    L, R = rods["length"], rods["radius"]
    b = -(1 + 2 * np.log(R / np.maximum(L, 1e-300)))
    sph = L < 2 * R
    rad = 0.5 * L + R
    zpara = np.where(sph, 6 * np.pi * rad * viscosity, 8 * np.pi * L * viscosity / (2 * b))
    zrot = np.where(sph, 8 * np.pi * rad**3 * viscosity, 2 * np.pi * viscosity * L**3 / (3 * (b + 2)))
    v = np.zeros((n, 6))
    v[:, :3] = rng.normal(size=(n, 3)) * np.sqrt(2 * kbt / (zpara * dt))[:, None]
    v[:, 3:] = rng.normal(size=(n, 3)) * np.sqrt(2 * kbt / (zrot * dt))[:, None]
    v[rods["immovable"] != 0] = 0
    return v.reshape(-1)


def read_rod_file(path):
    """Parse the reference's rod file format (SylinderSystem.cpp:317-344): two header lines, then
    `C|S gid radius mx my mz px py pz [group]`."""
    gid, rad, m, p, imm = [], [], [], [], []
    with open(path) as f:
        lines = f.readlines()[2:]
    for ln in lines:
        t = ln.split()
        if not t or t[0] not in ("C", "S"):
            continue
        imm.append(1 if t[0] == "S" else 0)
        gid.append(int(t[1]))
        rad.append(float(t[2]))
        m.append([float(x) for x in t[3:6]])
        p.append([float(x) for x in t[6:9]])
    m, p = np.array(m), np.array(p)
    pos = (m + p) * 0.5
    d = p - m
    length = np.sqrt((d**2).sum(axis=1))
    quat = quat_from_z_to(np.where(length[:, None] > 1e-7, d, np.array([0.0, 0.0, 1.0])))
    return dict(gid=np.array(gid, dtype=np.int32), pos=pos, quat=quat, length=length, radius=np.array(rad),
                immovable=np.array(imm, dtype=np.uint8))


def canonical_order(blocks):
    """sort key used everywhere for comparisons: (bilateral, gidI, gidJ, labJ) -- the reference's own
    Verify.py sorts the same way (Sylinder/Test2_MixLink/Verify.py:82-83)."""
    return np.lexsort((blocks["labJ"][:, 2], blocks["labJ"][:, 1], blocks["labJ"][:, 0], blocks["gidJ"],
                       blocks["gidI"], blocks["bilateral"]))
