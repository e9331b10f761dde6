"""Independent check of the MJCF loader's inertia inference ("the oracle and the CUDA path share the loader"):
body mass, centre of mass and principal moments of inertia inferred from geoms (sphere / capsule / box / ellipsoid, arbitrary
poses, several geoms per body, per-geom densities) against a Monte-Carlo volume integration that uses nothing of the loader --
the test writes the XML itself, so it knows every geom's parameters."""
import numpy as np

from brax_rodent_run_b200 import mjcf


def _rot(q):
    w, x, y, z = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def _inside(kind, size, p):
    if kind == "sphere":
        return (p ** 2).sum(1) <= size[0] ** 2
    if kind == "ellipsoid":
        return ((p / size) ** 2).sum(1) <= 1.0
    if kind == "box":
        return np.all(np.abs(p) <= size, axis=1)
    if kind == "capsule":  # axis z, radius size[0], half-length size[1]
        zc = np.clip(p[:, 2], -size[1], size[1])
        return p[:, 0] ** 2 + p[:, 1] ** 2 + (p[:, 2] - zc) ** 2 <= size[0] ** 2
    raise ValueError(kind)


def _mc_body(geoms, rng, n=400_000):
    """mass, com, inertia tensor about the com (body frame) of a list of (kind, size, pos, quat, density), geom by geom (MuJoCo
    adds overlapping geoms, it does not take the union)."""
    mass, first, second = 0.0, np.zeros(3), np.zeros((3, 3))
    for kind, size, pos, quat, rho in geoms:
        size = np.asarray(size, float)
        if kind == "sphere":
            half = np.full(3, size[0])
        elif kind == "capsule":
            half = np.array([size[0], size[0], size[1] + size[0]])
        else:
            half = size
        p = rng.uniform(-1, 1, (n, 3)) * half
        keep = _inside(kind, size, p)
        dm = rho * np.prod(2 * half) / n                 # mass per sample point
        pb = p[keep] @ _rot(np.asarray(quat, float)).T + np.asarray(pos, float)
        mass += dm * keep.sum()
        first += dm * pb.sum(0)
        second += dm * (pb.T @ pb)
    com = first / mass
    cov = second - mass * np.outer(com, com)             # sum m (r - c)(r - c)'
    return mass, com, np.trace(cov) * np.eye(3) - cov


BODIES = {
    "one_capsule": [("capsule", (0.02, 0.05), (0.01, -0.02, 0.03), (0.9, 0.1, 0.3, -0.2), 1500.0)],
    "sphere_and_box": [("sphere", (0.03,), (0.0, 0.0, 0.0), (1, 0, 0, 0), 468.25),
                       ("box", (0.02, 0.01, 0.04), (0.05, 0.01, -0.02), (0.7, 0.2, -0.5, 0.4), 2041.0)],
    "three_mixed": [("ellipsoid", (0.03, 0.015, 0.02), (-0.01, 0.02, 0.0), (0.6, -0.3, 0.2, 0.7), 1100.0),
                    ("capsule", (0.008, 0.0175), (0.0086, 0.0107, -0.0129), (0.8, 0.3, -0.35, 0.1), 1561.75),
                    ("capsule", (0.0032, 0.0124), (0.03, -0.01, 0.01), (0.5, 0.5, -0.5, 0.5), 2041.16)],
}


def _xml(bodies):
    out = ['<mujoco><compiler angle="radian"/><worldbody>']
    for i, (name, geoms) in enumerate(bodies.items()):
        out.append(f'<body name="{name}" pos="{0.3 * i} 0 0.5"><joint type="hinge" axis="0 1 0"/>')
        for kind, size, pos, quat, rho in geoms:
            out.append(f'<geom type="{kind}" size="{" ".join(map(str, size))}" pos="{" ".join(map(str, pos))}" '
                       f'quat="{" ".join(map(str, quat))}" density="{rho}" contype="0" conaffinity="0"/>')
        out.append("</body>")
    out.append("</worldbody></mujoco>")
    return "\n".join(out)


def test_inferred_inertia_matches_monte_carlo(tmp_path):
    path = tmp_path / "bodies.xml"
    path.write_text(_xml(BODIES))
    m = mjcf.load_xml(str(path))
    rng = np.random.default_rng(0)
    for b, (name, geoms) in enumerate(BODIES.items(), start=1):
        mass, com, inertia = _mc_body(geoms, rng)
        assert abs(m.body_mass[b] - mass) <= 5e-3 * mass, (name, m.body_mass[b], mass)
        extent = max(max(np.asarray(g[1], float)) for g in geoms)
        assert np.abs(m.body_ipos[b] - com).max() <= 5e-3 * extent, (name, m.body_ipos[b], com)
        # principal moments: the loader stores the eigenvalues (body_inertia) in the frame body_iquat
        want = np.sort(np.linalg.eigvalsh(inertia))
        got = np.sort(m.body_inertia[b])
        assert np.abs(got - want).max() <= 1e-2 * want.max(), (name, got, want)
        # and the frame: R diag(I) R' must reproduce the tensor, not only its spectrum
        R = _rot(m.body_iquat[b])
        assert np.abs(R @ np.diag(m.body_inertia[b]) @ R.T - inertia).max() <= 1.5e-2 * want.max(), name
