"""MJCF -> flat model tables (the "model compile" stage of the hot path).

Replaces, for the rodent model family only, what the reference obtains from
`mujoco.MjModel.from_xml_path` (Rodent_Env_Brax.py:41) followed by
`brax.io.mjcf.load_model` (Rodent_Env_Brax.py:51): MJCF parsing with nested
default classes, body inertia inference from geoms, kinematic-tree flattening,
`qpos0`/`qpos_spring`, the `mj_setConst` quantities (`dof_invweight0`,
`body_invweight0`, `stat.meaninertia`, `body_subtreemass`) and MJX's static
collision-pair table.  Everything is computed in float64 on the host, once;
the result is a `FlatModel` of numpy arrays that is serialised into the blob
consumed by both the CPU oracle (oracle/) and the CUDA library (csrc/).

Supported MJCF subset = what models/rodent_{0,new,optimized,pair}.xml use
(SURVEY.md Appendix A).  Anything else (tendons, equality constraints,
meshes, non-plane collision pairs ...) raises NotImplementedError, mirroring
Brax/MJX behaviour on unsupported features.
"""
from __future__ import annotations

import copy
import dataclasses
import math
import xml.etree.ElementTree as ET
from typing import Dict, List, Optional, Tuple

import numpy as np

# MuJoCo enums (mjtGeom, mjtJoint)
GEOM_PLANE, GEOM_HFIELD, GEOM_SPHERE, GEOM_CAPSULE, GEOM_ELLIPSOID, GEOM_CYLINDER, GEOM_BOX, GEOM_MESH = range(8)
_GEOM_TYPES = {"plane": 0, "hfield": 1, "sphere": 2, "capsule": 3, "ellipsoid": 4, "cylinder": 5, "box": 6, "mesh": 7}
JNT_FREE, JNT_BALL, JNT_SLIDE, JNT_HINGE = range(4)
MJ_MINVAL = 1e-15

# collision function ids used in the static pair table
PAIR_PLANE_SPHERE, PAIR_PLANE_CAPSULE, PAIR_PLANE_ELLIPSOID = 0, 1, 2
_PAIR_NCON = {PAIR_PLANE_SPHERE: 1, PAIR_PLANE_CAPSULE: 2, PAIR_PLANE_ELLIPSOID: 1}


# --------------------------------------------------------------------------- math helpers (float64)
def _vec(s, n=None, default=None):
    if s is None:
        return None if default is None else np.array(default, dtype=np.float64)
    v = np.array([float(x) for x in s.split()], dtype=np.float64)
    if n is not None and v.size < n and default is not None:
        d = np.array(default, dtype=np.float64)
        d[: v.size] = v
        v = d
    return v


def quat_mul(a, b):
    return np.array([
        a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
        a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
        a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
        a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0],
    ])


def quat_to_mat(q):
    w, x, y, z = q
    return np.array([
        [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z],
    ])


def mat_to_quat(m):
    # Shepperd's method, returns w>=0 normalised quaternion
    t = np.trace(m)
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s])
    elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
        s = math.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
        q = np.array([(m[2, 1] - m[1, 2]) / s, 0.25 * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s])
    elif m[1, 1] > m[2, 2]:
        s = math.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
        q = np.array([(m[0, 2] - m[2, 0]) / s, (m[0, 1] + m[1, 0]) / s, 0.25 * s, (m[1, 2] + m[2, 1]) / s])
    else:
        s = math.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
        q = np.array([(m[1, 0] - m[0, 1]) / s, (m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, 0.25 * s])
    q = q / np.linalg.norm(q)
    return q if q[0] >= 0 else -q


def axis_angle_quat(axis, angle):
    s = math.sin(angle * 0.5)
    return np.array([math.cos(angle * 0.5), axis[0] * s, axis[1] * s, axis[2] * s])


def euler_to_quat(e, seq="xyz"):
    """MuJoCo eulerseq: lower-case = intrinsic (rotating) axes, applied left to right."""
    q = np.array([1.0, 0, 0, 0])
    for ang, ax in zip(e, seq):
        a = np.zeros(3)
        a["xyz".index(ax.lower())] = 1.0
        r = axis_angle_quat(a, ang)
        q = quat_mul(q, r) if ax.islower() else quat_mul(r, q)
    return q


def zaxis_to_quat(z):
    """Minimal rotation taking (0,0,1) to z (mjuu_z2quat)."""
    z = np.asarray(z, dtype=np.float64)
    z = z / np.linalg.norm(z)
    axis = np.cross([0.0, 0.0, 1.0], z)
    s = np.linalg.norm(axis)
    if s < 1e-10:
        return np.array([1.0, 0, 0, 0]) if z[2] > 0 else np.array([0.0, 1.0, 0, 0])
    axis /= s
    ang = math.atan2(s, z[2])
    return axis_angle_quat(axis, ang)


# --------------------------------------------------------------------------- parsed (pre-compile) objects
@dataclasses.dataclass
class _Body:
    name: str
    parent: int
    pos: np.ndarray
    quat: np.ndarray
    joints: list = dataclasses.field(default_factory=list)
    geoms: list = dataclasses.field(default_factory=list)


_ORIENT_KEYS = ("quat", "euler", "zaxis", "axisangle", "xyaxes")


class _Defaults:
    """Nested <default class=...> resolution.  `get(tag, cls)` returns the merged attribute dict."""

    def __init__(self, root: Optional[ET.Element]):
        self.cls: Dict[str, Dict[str, Dict[str, str]]] = {"main": {}}
        if root is not None:
            self._walk(root, "main", {})

    def _walk(self, node, name, inherited):
        mine = {k: dict(v) for k, v in inherited.items()}
        for child in node:
            if child.tag == "default":
                continue
            d = mine.setdefault(child.tag, {})
            # an orientation given at this level replaces any inherited orientation spec
            if any(k in child.attrib for k in _ORIENT_KEYS):
                for k in _ORIENT_KEYS:
                    d.pop(k, None)
            d.update(child.attrib)
        self.cls[name] = mine
        for child in node:
            if child.tag == "default":
                cname = child.attrib.get("class")
                if cname is None:
                    raise ValueError("nested <default> requires a class attribute")
                self._walk(child, cname, mine)

    def resolve(self, elem: ET.Element, childclass: Optional[str], tag: Optional[str] = None, use_defaults=True):
        tag = tag or elem.tag
        out: Dict[str, str] = {}
        if use_defaults:
            cname = elem.attrib.get("class", childclass or "main")
            if cname not in self.cls:
                raise ValueError(f"unknown default class '{cname}'")
            out.update(self.cls[cname].get(tag, {}))
        if any(k in elem.attrib for k in _ORIENT_KEYS):
            for k in _ORIENT_KEYS:
                out.pop(k, None)
        out.update(elem.attrib)
        return out


def _orientation(attr: Dict[str, str], eulerseq: str, angle_scale: float) -> np.ndarray:
    if "quat" in attr:
        q = _vec(attr["quat"])
        return q / np.linalg.norm(q)
    if "euler" in attr:
        return euler_to_quat(_vec(attr["euler"]) * angle_scale, eulerseq)
    if "zaxis" in attr:
        return zaxis_to_quat(_vec(attr["zaxis"]))
    if "axisangle" in attr:
        v = _vec(attr["axisangle"])
        return axis_angle_quat(v[:3] / np.linalg.norm(v[:3]), v[3] * angle_scale)
    if "xyaxes" in attr:
        v = _vec(attr["xyaxes"])
        x = v[:3] / np.linalg.norm(v[:3])
        y = v[3:] - x * np.dot(x, v[3:])
        y /= np.linalg.norm(y)
        return mat_to_quat(np.stack([x, y, np.cross(x, y)], axis=1))
    return np.array([1.0, 0, 0, 0])


def _bool(s, default=False):
    if s is None:
        return default
    return s.strip().lower() == "true"


# --------------------------------------------------------------------------- the flat model
@dataclasses.dataclass
class FlatModel:
    """Flattened model tables (all float arrays float64, all index arrays int32).

    Field names follow MuJoCo's mjModel where a counterpart exists."""

    # sizes
    nq: int = 0
    nv: int = 0
    nu: int = 0
    na: int = 0
    nbody: int = 0
    njnt: int = 0
    ngeom: int = 0
    nM: int = 0
    npair: int = 0
    ncon: int = 0
    nlimit: int = 0
    nefc: int = 0
    # options (mjOption)
    timestep: float = 0.002
    gravity: np.ndarray = dataclasses.field(default_factory=lambda: np.array([0, 0, -9.81]))
    tolerance: float = 1e-8
    ls_tolerance: float = 0.01
    impratio: float = 1.0
    solver: int = 1  # 0 = CG, 1 = Newton  (rr convention; MuJoCo default is Newton)
    iterations: int = 100
    ls_iterations: int = 50
    meaninertia: float = 1.0
    names: dict = dataclasses.field(default_factory=dict)
    arrays: Dict[str, np.ndarray] = dataclasses.field(default_factory=dict)

    def __getattr__(self, k):
        arrays = self.__dict__.get("arrays", {})
        if k in arrays:
            return arrays[k]
        raise AttributeError(k)

    def replace_options(self, **kw) -> "FlatModel":
        m = copy.copy(self)
        m.arrays = dict(self.arrays)
        for k, v in kw.items():
            if not hasattr(m, k):
                raise AttributeError(k)
            setattr(m, k, v)
        return m

    # ---- (de)serialisation of compiled models (assets/*.npz) -------------------------------------
    _SCALARS = ("nq", "nv", "nu", "na", "nbody", "njnt", "ngeom", "nM", "npair", "ncon", "nlimit", "nefc",
                "timestep", "tolerance", "ls_tolerance", "impratio", "solver", "iterations", "ls_iterations",
                "meaninertia")

    def save(self, path):
        d = {f"arr__{k}": v for k, v in self.arrays.items()}
        for k in self._SCALARS:
            d[f"sc__{k}"] = np.array(getattr(self, k))
        d["sc__gravity"] = np.asarray(self.gravity, dtype=np.float64)
        for k, v in self.names.items():
            d[f"names__{k}"] = np.array(v)
        np.savez_compressed(path, **d)

    @classmethod
    def load(cls, path) -> "FlatModel":
        z = np.load(path, allow_pickle=False)
        m = cls()
        for k in z.files:
            if k.startswith("arr__"):
                m.arrays[k[5:]] = z[k]
            elif k.startswith("names__"):
                m.names[k[7:]] = [str(s) for s in z[k]]
            elif k == "sc__gravity":
                m.gravity = z[k].astype(np.float64)
            elif k.startswith("sc__"):
                cur = getattr(m, k[4:])
                setattr(m, k[4:], type(cur)(z[k].item()))
        return m


# --------------------------------------------------------------------------- parser
def _expand_replicates(node: ET.Element, eulerseq: str, angle_scale: float):
    """Expand <replicate count offset euler sep> in place (MuJoCo >= 3.1.4, models/rodent_pair.xml:163).

    Copy i is wrapped in a frame rotated by i*euler and translated by i*offset (accumulated); names get
    `sep + i` appended.  Returns the list of (old_name -> [new names]) maps for replicated elements so
    that actuators/sensors referencing them can be replicated as well."""
    renames: List[Dict[str, List[str]]] = []
    changed = True
    while changed:
        changed = False
        for parent in list(node.iter()):
            for idx, child in enumerate(list(parent)):
                if child.tag != "replicate":
                    continue
                count = int(child.attrib["count"])
                sep = child.attrib.get("sep", "")
                offset = _vec(child.attrib.get("offset"), 3, [0, 0, 0])
                euler = _vec(child.attrib.get("euler"), 3, [0, 0, 0]) * angle_scale
                dq = euler_to_quat(euler, eulerseq)
                ren: Dict[tuple, List[str]] = {}
                new_children = []
                pos = np.zeros(3)
                quat = np.array([1.0, 0, 0, 0])
                for i in range(count):
                    for sub in child:
                        c = copy.deepcopy(sub)
                        for e in c.iter():
                            if "name" in e.attrib:
                                old = e.attrib["name"]
                                e.attrib["name"] = f"{old}{sep}{i}"
                                kind = "joint" if e.tag == "freejoint" else e.tag
                                ren.setdefault((kind, old), []).append(e.attrib["name"])
                        # apply the frame of copy i to the top-level element
                        if c.tag in ("body", "geom", "site", "camera", "light"):
                            a = dict(c.attrib)
                            p = _vec(a.get("pos"), 3, [0, 0, 0])
                            q = _orientation(a, eulerseq, angle_scale)
                            for k in _ORIENT_KEYS:
                                c.attrib.pop(k, None)
                            np_ = pos + quat_to_mat(quat) @ p
                            nq_ = quat_mul(quat, q)
                            c.attrib["pos"] = " ".join(repr(float(x)) for x in np_)
                            c.attrib["quat"] = " ".join(repr(float(x)) for x in nq_)
                        new_children.append(c)
                    pos = pos + quat_to_mat(quat) @ offset
                    quat = quat_mul(quat, dq)
                parent.remove(child)
                for k, c in enumerate(new_children):
                    parent.insert(idx + k, c)
                renames.append(ren)
                changed = True
                break
            if changed:
                break
    return renames


def _replicate_referrers(section: Optional[ET.Element], renames, ref_attrs):
    """Replicate actuators / sensors whose target was replicated (MuJoCo does this automatically)."""
    if section is None:
        return
    for ren in renames:
        for elem in list(section):
            targets = [a for a in ref_attrs if a in elem.attrib and (a, elem.attrib[a]) in ren]
            if not targets:
                continue
            n = len(ren[(targets[0], elem.attrib[targets[0]])])
            idx = list(section).index(elem)
            section.remove(elem)
            for i in range(n):
                c = copy.deepcopy(elem)
                for a in targets:
                    c.attrib[a] = ren[(a, elem.attrib[a])][i]
                if "name" in c.attrib:
                    new_target = c.attrib[targets[0]]
                    old_target = elem.attrib[targets[0]]
                    c.attrib["name"] = c.attrib["name"] + new_target[len(old_target):]
                section.insert(idx + i, c)


def _rescale_subtree(node: ET.Element, position_factor: float, size_factor: float) -> None:
    """dm_control.locomotion.walkers.rescale.rescale_subtree on the XML tree (preprocessing/mjx_preprocess.py:80-84):
    every explicit `pos` under the body tree (bodies, geoms, joints, sites, cameras, inertials) is multiplied by
    `position_factor`, every explicit `size` by `size_factor`, `fromto` about its scaled midpoint; defaults are untouched."""
    fmt = lambda v: " ".join(repr(float(x)) for x in v)
    for child in list(node):
        a = child.attrib
        if "fromto" in a:
            ft = _vec(a["fromto"], 6)
            mid, half = position_factor * 0.5 * (ft[3:] + ft[:3]), size_factor * 0.5 * (ft[3:] - ft[:3])
            a["fromto"] = fmt(np.concatenate([mid - half, mid + half]))
        if "pos" in a:
            a["pos"] = fmt(position_factor * _vec(a["pos"]))
        if "size" in a:
            a["size"] = fmt(size_factor * _vec(a["size"]))
        if child.tag in ("body", "worldbody"):
            _rescale_subtree(child, position_factor, size_factor)


def load_xml(path: str, rescale: Optional[Tuple[float, float]] = None) -> FlatModel:
    """Parse + compile an MJCF file into a FlatModel.  `rescale=(position_factor, size_factor)` applies dm_control's
    rescale_subtree to the body tree first (the reference's mocap preprocessing uses 0.9, 0.9)."""
    try:
        tree = ET.parse(path)
    except (ET.ParseError, OSError) as e:  # MuJoCo raises ValueError on bad XML / missing file
        raise ValueError(f"XML Error: {e}") from e
    root = tree.getroot()
    if root.tag != "mujoco":
        raise ValueError("XML Error: root element must be <mujoco>")

    # ---- compiler / option ---------------------------------------------------------------------
    angle_scale = math.pi / 180.0  # MuJoCo default: degrees
    eulerseq = "xyz"
    for comp in root.findall("compiler"):
        if "angle" in comp.attrib:
            angle_scale = 1.0 if comp.attrib["angle"] == "radian" else math.pi / 180.0
        eulerseq = comp.attrib.get("eulerseq", eulerseq)
        for bad in ("inertiafromgeom", "settotalmass", "balanceinertia", "fusestatic", "coordinate"):
            if bad in comp.attrib:
                raise NotImplementedError(f"<compiler {bad}=...> is not supported")
    m = FlatModel()
    for opt in root.findall("option"):
        a = opt.attrib
        m.timestep = float(a.get("timestep", m.timestep))
        m.gravity = _vec(a.get("gravity"), 3, m.gravity)
        m.tolerance = float(a.get("tolerance", m.tolerance))
        m.ls_tolerance = float(a.get("ls_tolerance", m.ls_tolerance))
        m.impratio = float(a.get("impratio", m.impratio))
        m.iterations = int(a.get("iterations", m.iterations))
        m.ls_iterations = int(a.get("ls_iterations", m.ls_iterations))
        if "solver" in a:
            m.solver = {"cg": 0, "newton": 1}[a["solver"].lower()]
        if a.get("cone", "pyramidal") != "pyramidal":
            raise NotImplementedError("only pyramidal friction cones are supported (MJX 3.1.x)")
        if a.get("integrator", "Euler") != "Euler":
            raise NotImplementedError("only the Euler integrator is supported")

    for tag in ("equality", "keyframe", "custom", "extension", "deformable"):
        sec = root.find(tag)
        if sec is not None and len(sec):
            raise NotImplementedError(f"<{tag}> is not supported")
    tendon = root.find("tendon")
    if tendon is not None and len(tendon):
        raise NotImplementedError("tendons are not supported (MJX/Brax reject models/rodent_cpu.xml as well)")

    defaults = _Defaults(root.find("default"))
    worldbody = root.find("worldbody")
    if worldbody is None:
        raise ValueError("XML Error: missing <worldbody>")
    if rescale is not None:
        _rescale_subtree(worldbody, float(rescale[0]), float(rescale[1]))
    renames = _expand_replicates(worldbody, eulerseq, angle_scale)
    _replicate_referrers(root.find("actuator"), renames, ("joint",))
    _replicate_referrers(root.find("sensor"), renames, ("site", "body", "joint"))

    # ---- walk the body tree (depth-first pre-order = MuJoCo body numbering) ---------------------
    bodies: List[_Body] = [_Body("world", 0, np.zeros(3), np.array([1.0, 0, 0, 0]))]

    def parse_geom(elem, childclass):
        a = defaults.resolve(elem, childclass, "geom")
        gtype = _GEOM_TYPES[a.get("type", "sphere")]
        if gtype in (GEOM_MESH, GEOM_HFIELD):
            raise NotImplementedError("mesh / hfield geoms are not supported")
        if "fromto" in a:
            raise NotImplementedError("geom fromto is not supported")
        if "mass" in a:
            raise NotImplementedError("geom mass attribute is not supported")
        size = _vec(a.get("size"), 3, [0, 0, 0])
        return dict(
            name=a.get("name", ""), type=gtype, size=size,
            pos=_vec(a.get("pos"), 3, [0, 0, 0]), quat=_orientation(a, eulerseq, angle_scale),
            density=float(a.get("density", 1000.0)),
            contype=int(a.get("contype", 1)), conaffinity=int(a.get("conaffinity", 1)),
            condim=int(a.get("condim", 3)), priority=int(a.get("priority", 0)),
            friction=_vec(a.get("friction"), 3, [1.0, 0.005, 0.0001]),
            solref=_vec(a.get("solref"), 2, [0.02, 1.0]),
            solimp=_vec(a.get("solimp"), 5, [0.9, 0.95, 0.001, 0.5, 2.0]),
            solmix=float(a.get("solmix", 1.0)), margin=float(a.get("margin", 0.0)), gap=float(a.get("gap", 0.0)),
        )

    def parse_joint(elem, childclass):
        if elem.tag == "freejoint":
            a = dict(elem.attrib)  # <freejoint> takes no defaults
            jtype = JNT_FREE
            a_def = {}
        else:
            a = defaults.resolve(elem, childclass, "joint")
            jtype = {"free": JNT_FREE, "ball": JNT_BALL, "slide": JNT_SLIDE, "hinge": JNT_HINGE}[a.get("type", "hinge")]
            a_def = a
        if jtype in (JNT_BALL, JNT_SLIDE):
            raise NotImplementedError("ball / slide joints are not supported")
        axis = _vec(a_def.get("axis"), 3, [0, 0, 1])
        axis = axis / np.linalg.norm(axis)
        rng = _vec(a_def.get("range"), 2, [0, 0])
        lim = a_def.get("limited", "auto")
        limited = (lim == "true") or (lim == "auto" and "range" in a_def and rng[0] < rng[1])
        if jtype == JNT_FREE:
            limited = False
        if float(a_def.get("frictionloss", 0.0)) != 0.0:
            raise NotImplementedError("joint frictionloss is not supported")
        return dict(
            name=a.get("name", ""), type=jtype, pos=_vec(a_def.get("pos"), 3, [0, 0, 0]), axis=axis,
            range=rng * (angle_scale if jtype == JNT_HINGE else 1.0), limited=limited,
            stiffness=float(a_def.get("stiffness", 0.0)), damping=float(a_def.get("damping", 0.0)),
            armature=float(a_def.get("armature", 0.0)), margin=float(a_def.get("margin", 0.0)),
            ref=float(a_def.get("ref", 0.0)) * (angle_scale if jtype == JNT_HINGE else 1.0),
            springref=float(a_def.get("springref", 0.0)) * (angle_scale if jtype == JNT_HINGE else 1.0),
            solref=_vec(a_def.get("solreflimit"), 2, [0.02, 1.0]),
            solimp=_vec(a_def.get("solimplimit"), 5, [0.9, 0.95, 0.001, 0.5, 2.0]),
        )

    def walk(elem, parent_id, childclass):
        for child in elem:
            if child.tag == "geom":
                bodies[parent_id].geoms.append(parse_geom(child, childclass))
            elif child.tag in ("joint", "freejoint"):
                bodies[parent_id].joints.append(parse_joint(child, childclass))
            elif child.tag == "inertial":
                raise NotImplementedError("explicit <inertial> is not supported")
            elif child.tag == "frame":
                raise NotImplementedError("<frame> is not supported")
        for child in elem:
            if child.tag == "body":
                a = child.attrib
                b = _Body(a.get("name", ""), parent_id, _vec(a.get("pos"), 3, [0, 0, 0]),
                          _orientation(a, eulerseq, angle_scale))
                bodies.append(b)
                bid = len(bodies) - 1
                walk(child, bid, a.get("childclass", childclass))

    # MuJoCo numbers bodies depth-first, but a body's own geoms/joints are collected before recursing.
    walk(worldbody, 0, None)
    return _compile(m, bodies, root, defaults)


# --------------------------------------------------------------------------- compile
def _geom_mass_inertia(g):
    t, s, rho = g["type"], g["size"], g["density"]
    if t == GEOM_PLANE:
        return 0.0, np.zeros(3)
    if t == GEOM_SPHERE:
        r = s[0]
        mass = rho * 4.0 / 3.0 * math.pi * r ** 3
        return mass, np.full(3, 0.4 * mass * r * r)
    if t == GEOM_CAPSULE:
        r, h = s[0], 2.0 * s[1]
        mass = rho * (math.pi * r * r * h + 4.0 / 3.0 * math.pi * r ** 3)
        sm = mass * 4 * r / (4 * r + 3 * h)
        cm = mass - sm
        ix = cm * (3 * r * r + h * h) / 12.0
        iz = cm * r * r / 2.0
        si = 2 * sm * r * r / 5.0
        ix += si + sm * h * (3 * r + 2 * h) / 8.0
        iz += si
        return mass, np.array([ix, ix, iz])
    if t == GEOM_ELLIPSOID:
        mass = rho * 4.0 / 3.0 * math.pi * s[0] * s[1] * s[2]
        return mass, mass / 5.0 * np.array([s[1] ** 2 + s[2] ** 2, s[0] ** 2 + s[2] ** 2, s[0] ** 2 + s[1] ** 2])
    if t == GEOM_BOX:
        mass = rho * 8.0 * s[0] * s[1] * s[2]
        return mass, mass / 3.0 * np.array([s[1] ** 2 + s[2] ** 2, s[0] ** 2 + s[2] ** 2, s[0] ** 2 + s[1] ** 2])
    if t == GEOM_CYLINDER:
        r, h = s[0], 2.0 * s[1]
        mass = rho * math.pi * r * r * h
        ix = mass * (3 * r * r + h * h) / 12.0
        return mass, np.array([ix, ix, mass * r * r / 2.0])
    raise NotImplementedError(f"geom type {t}")


def _compile(m: FlatModel, bodies: List[_Body], root: ET.Element, defaults: _Defaults) -> FlatModel:
    A: Dict[str, np.ndarray] = {}
    nbody = len(bodies)
    i32 = np.int32

    # ---- bodies, joints, dofs ------------------------------------------------------------------
    body_parentid = np.array([b.parent for b in bodies], dtype=i32)
    body_rootid = np.zeros(nbody, dtype=i32)
    for b in range(1, nbody):
        body_rootid[b] = b if body_parentid[b] == 0 else body_rootid[body_parentid[b]]
    body_pos = np.stack([b.pos for b in bodies])
    body_quat = np.stack([b.quat for b in bodies])
    body_jntadr = np.full(nbody, -1, dtype=i32)
    body_jntnum = np.zeros(nbody, dtype=i32)
    body_dofadr = np.full(nbody, -1, dtype=i32)
    body_dofnum = np.zeros(nbody, dtype=i32)
    jnts, jnt_bodyid = [], []
    qadr = dadr = 0
    jnt_qposadr, jnt_dofadr = [], []
    dof_bodyid, dof_jntid, dof_parentid = [], [], []
    last_dof_of_body = np.full(nbody, -1, dtype=i32)  # last dof on the path root->body (inclusive)
    for b, body in enumerate(bodies):
        par_last = last_dof_of_body[body.parent] if b > 0 else -1
        last = par_last
        if body.joints:
            body_jntadr[b] = len(jnts)
            body_jntnum[b] = len(body.joints)
            body_dofadr[b] = dadr
        for j in body.joints:
            if j["type"] == JNT_FREE and (body.parent != 0 or len(body.joints) != 1):
                raise ValueError("free joint can only be used on top level, as the only joint")
            jnts.append(j)
            jnt_bodyid.append(b)
            jnt_qposadr.append(qadr)
            jnt_dofadr.append(dadr)
            nd = 6 if j["type"] == JNT_FREE else 1
            for k in range(nd):
                dof_bodyid.append(b)
                dof_jntid.append(len(jnts) - 1)
                dof_parentid.append(last)
                last = dadr + k
            qadr += 7 if j["type"] == JNT_FREE else 1
            dadr += nd
        body_dofnum[b] = dadr - body_dofadr[b] if body.joints else 0
        last_dof_of_body[b] = last
    nq, nv, njnt = qadr, dadr, len(jnts)
    dof_parentid = np.array(dof_parentid, dtype=i32)
    dof_bodyid = np.array(dof_bodyid, dtype=i32)
    dof_jntid = np.array(dof_jntid, dtype=i32)

    # ---- body inertial frames from geoms -------------------------------------------------------
    body_mass = np.zeros(nbody)
    body_ipos = np.zeros((nbody, 3))
    body_iquat = np.tile(np.array([1.0, 0, 0, 0]), (nbody, 1))
    body_inertia = np.zeros((nbody, 3))
    geoms, geom_bodyid = [], []
    for b, body in enumerate(bodies):
        gm = []
        for g in body.geoms:
            geoms.append(g)
            geom_bodyid.append(b)
            mass, inr = _geom_mass_inertia(g)
            gm.append((mass, inr, g))
        tot = sum(x[0] for x in gm)
        body_mass[b] = tot
        if tot <= 0:
            continue  # massless frame body (e.g. the `walker` wrapper of rodent_new.xml:161); checked below
        com = sum(x[0] * x[2]["pos"] for x in gm) / tot
        I = np.zeros((3, 3))
        for mass, inr, g in gm:
            R = quat_to_mat(g["quat"])
            d = g["pos"] - com
            I += R @ np.diag(inr) @ R.T + mass * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
        w, V = np.linalg.eigh(0.5 * (I + I.T))
        order = np.argsort(-w)  # MuJoCo: principal moments in decreasing order
        w, V = w[order], V[:, order]
        if np.linalg.det(V) < 0:
            V[:, 2] = -V[:, 2]
        body_ipos[b] = com
        body_iquat[b] = mat_to_quat(V)
        body_inertia[b] = w
    ngeom = len(geoms)

    body_subtreemass = body_mass.copy()
    for b in range(nbody - 1, 0, -1):
        body_subtreemass[body_parentid[b]] += body_subtreemass[b]
    for b in range(1, nbody):
        if bodies[b].joints and body_subtreemass[b] < MJ_MINVAL:
            raise ValueError(f"mass and inertia of moving bodies must be larger than mjMINVAL: {bodies[b].name}")
    # subtree extents (DFS pre-order => subtree(b) = [b, b + body_subtreesize[b]))
    body_subtreesize = np.ones(nbody, dtype=i32)
    for b in range(nbody - 1, 0, -1):
        body_subtreesize[body_parentid[b]] += body_subtreesize[b]

    # ---- joint / dof parameter arrays ----------------------------------------------------------
    qpos0 = np.zeros(nq)
    qpos_spring = np.zeros(nq)
    for j, jn in enumerate(jnts):
        a = jnt_qposadr[j]
        if jn["type"] == JNT_FREE:
            b = jnt_bodyid[j]
            qpos0[a:a + 3] = body_pos[b]
            qpos0[a + 3:a + 7] = body_quat[b]
            qpos_spring[a:a + 7] = qpos0[a:a + 7]
        else:
            qpos0[a] = jn["ref"]
            qpos_spring[a] = jn["springref"]
    dof_armature = np.array([jnts[j]["armature"] for j in dof_jntid])
    dof_damping = np.array([jnts[j]["damping"] for j in dof_jntid])

    A.update(
        body_parentid=body_parentid, body_rootid=body_rootid, body_jntadr=body_jntadr, body_jntnum=body_jntnum,
        body_dofadr=body_dofadr, body_dofnum=body_dofnum, body_pos=body_pos, body_quat=body_quat,
        body_ipos=body_ipos, body_iquat=body_iquat, body_inertia=body_inertia, body_mass=body_mass,
        body_subtreemass=body_subtreemass, body_subtreesize=body_subtreesize,
        body_lastdof=last_dof_of_body.astype(i32),
        jnt_type=np.array([j["type"] for j in jnts], dtype=i32), jnt_bodyid=np.array(jnt_bodyid, dtype=i32),
        jnt_qposadr=np.array(jnt_qposadr, dtype=i32), jnt_dofadr=np.array(jnt_dofadr, dtype=i32),
        jnt_pos=np.stack([j["pos"] for j in jnts]), jnt_axis=np.stack([j["axis"] for j in jnts]),
        jnt_range=np.stack([j["range"] for j in jnts]), jnt_limited=np.array([j["limited"] for j in jnts], dtype=i32),
        jnt_stiffness=np.array([j["stiffness"] for j in jnts]), jnt_margin=np.array([j["margin"] for j in jnts]),
        jnt_solref=np.stack([j["solref"] for j in jnts]), jnt_solimp=np.stack([j["solimp"] for j in jnts]),
        qpos0=qpos0, qpos_spring=qpos_spring,
        dof_bodyid=dof_bodyid, dof_jntid=dof_jntid, dof_parentid=dof_parentid,
        dof_armature=dof_armature, dof_damping=dof_damping,
    )

    # ---- sparse mass-matrix layout: row i = ancestors of i in ascending order, diagonal last ----
    rowadr = np.zeros(nv, dtype=i32)
    rownnz = np.zeros(nv, dtype=i32)
    colind = []
    for i in range(nv):
        anc = []
        k = i
        while k >= 0:
            anc.append(k)
            k = dof_parentid[k]
        anc.reverse()
        rowadr[i] = len(colind)
        rownnz[i] = len(anc)
        colind.extend(anc)
    A.update(M_rowadr=rowadr, M_rownnz=rownnz, M_colind=np.array(colind, dtype=i32))
    nM = len(colind)

    # ---- geoms ---------------------------------------------------------------------------------
    def garr(key, dtype=np.float64):
        return np.array([g[key] for g in geoms], dtype=dtype) if geoms else np.zeros((0,), dtype=dtype)

    A.update(
        geom_type=garr("type", i32), geom_bodyid=np.array(geom_bodyid, dtype=i32), geom_pos=garr("pos"),
        geom_quat=garr("quat"), geom_size=garr("size"), geom_contype=garr("contype", i32),
        geom_conaffinity=garr("conaffinity", i32), geom_condim=garr("condim", i32), geom_priority=garr("priority", i32),
        geom_friction=garr("friction"), geom_solref=garr("solref"), geom_solimp=garr("solimp"),
        geom_solmix=garr("solmix"), geom_margin=garr("margin"), geom_gap=garr("gap"),
    )

    # ---- actuators -----------------------------------------------------------------------------
    jname = {j["name"]: k for k, j in enumerate(jnts) if j["name"]}
    acts = []
    asec = root.find("actuator")
    for elem in (asec if asec is not None else []):
        if elem.tag != "general":
            raise NotImplementedError(f"actuator shortcut <{elem.tag}> is not supported")
        a = defaults.resolve(elem, None, "general")
        if "joint" not in a:
            raise NotImplementedError("only joint transmissions are supported")
        if a["joint"] not in jname:
            raise ValueError(f"unknown joint '{a['joint']}' in actuator")
        j = jname[a["joint"]]
        if jnts[j]["type"] != JNT_HINGE:
            raise NotImplementedError("actuators on free joints are not supported")
        dyntype = {"none": 0, "integrator": 1, "filter": 2, "filterexact": 3}[a.get("dyntype", "none")]
        if dyntype not in (0, 2):
            raise NotImplementedError("dyntype must be none or filter")
        gaintype = {"fixed": 0, "affine": 1}[a.get("gaintype", "fixed")]
        biastype = {"none": 0, "affine": 1}[a.get("biastype", "none")]
        ctrlrange = _vec(a.get("ctrlrange"), 2, [0, 0])
        forcerange = _vec(a.get("forcerange"), 2, [0, 0])

        def lim(key, rng):
            v = a.get(key, "auto")
            return v == "true" or (v == "auto" and rng[0] < rng[1])

        acts.append(dict(
            name=a.get("name", ""), jnt=j, gear=_vec(a.get("gear"), 6, [1, 0, 0, 0, 0, 0])[0], dyntype=dyntype,
            dynprm=_vec(a.get("dynprm"), 3, [1, 0, 0])[0], gaintype=gaintype, biastype=biastype,
            gainprm=_vec(a.get("gainprm"), 3, [1, 0, 0]), biasprm=_vec(a.get("biasprm"), 3, [0, 0, 0]),
            ctrllimited=lim("ctrllimited", ctrlrange), ctrlrange=ctrlrange,
            forcelimited=lim("forcelimited", forcerange), forcerange=forcerange,
        ))
    nu = len(acts)
    na = sum(1 for a in acts if a["dyntype"] != 0)
    actadr, k = [], 0
    for a in acts:
        if a["dyntype"] != 0:
            actadr.append(k)
            k += 1
        else:
            actadr.append(-1)
    if na and any(x < 0 for x in actadr[actadr.index(0):]) if 0 in actadr else False:
        raise ValueError("stateless actuators must come before stateful ones")

    def aarr(key, dtype=np.float64, shape=()):
        return np.array([a[key] for a in acts], dtype=dtype) if acts else np.zeros((0,) + shape, dtype=dtype)

    A.update(
        actuator_jntid=aarr("jnt", i32), actuator_gear=aarr("gear"), actuator_dyntype=aarr("dyntype", i32),
        actuator_dynprm=aarr("dynprm"), actuator_gaintype=aarr("gaintype", i32), actuator_biastype=aarr("biastype", i32),
        actuator_gainprm=aarr("gainprm", shape=(3,)), actuator_biasprm=aarr("biasprm", shape=(3,)),
        actuator_ctrllimited=aarr("ctrllimited", i32), actuator_ctrlrange=aarr("ctrlrange", shape=(2,)),
        actuator_forcelimited=aarr("forcelimited", i32), actuator_forcerange=aarr("forcerange", shape=(2,)),
        actuator_actadr=np.array(actadr, dtype=i32),
    )

    # ---- static collision-pair table (MJX collision_driver: grouped by function, first-seen order) ----
    csec = root.find("contact")
    excludes = set()
    if csec is not None:
        bname = {b.name: k for k, b in enumerate(bodies)}
        for e in csec:
            if e.tag == "exclude":
                excludes.add(frozenset((bname[e.attrib["body1"]], bname[e.attrib["body2"]])))
            elif e.tag == "pair":
                raise NotImplementedError("explicit contact <pair> is not supported")
    body_weldid = np.zeros(nbody, dtype=i32)
    for b in range(1, nbody):
        body_weldid[b] = b if bodies[b].joints else body_weldid[body_parentid[b]]
    groups: Dict[int, list] = {}
    for g1 in range(ngeom):
        for g2 in range(g1 + 1, ngeom):
            a, b = geoms[g1], geoms[g2]
            if not ((a["contype"] & b["conaffinity"]) or (b["contype"] & a["conaffinity"])):
                continue
            b1, b2 = geom_bodyid[g1], geom_bodyid[g2]
            w1, w2 = body_weldid[b1], body_weldid[b2]
            if w1 == w2:
                continue
            if frozenset((b1, b2)) in excludes:
                continue
            # filterparent: skip parent-child weld pairs unless the parent is the world
            p1, p2 = body_weldid[body_parentid[w1]], body_weldid[body_parentid[w2]]
            if (w1 != 0 and w2 != 0) and (p1 == w2 or p2 == w1):
                continue
            t1, t2 = a["type"], b["type"]
            ga, gb = g1, g2
            if t1 > t2:  # MuJoCo orders the pair so that type1 <= type2
                t1, t2, ga, gb = t2, t1, g2, g1
            fn = {(GEOM_PLANE, GEOM_SPHERE): PAIR_PLANE_SPHERE, (GEOM_PLANE, GEOM_CAPSULE): PAIR_PLANE_CAPSULE,
                  (GEOM_PLANE, GEOM_ELLIPSOID): PAIR_PLANE_ELLIPSOID}.get((t1, t2))
            if fn is None:
                raise NotImplementedError(
                    f"collision between geom types {t1} and {t2} ({geoms[ga]['name']}, {geoms[gb]['name']}) is not supported")
            groups.setdefault(fn, []).append((ga, gb))
    pair_fn, pair_g1, pair_g2, pair_conadr = [], [], [], []
    pair_mu, pair_solref, pair_solimp, pair_margin = [], [], [], []
    ncon = 0
    for fn, plist in groups.items():  # dict preserves first-seen order
        for ga, gb in plist:
            a, b = geoms[ga], geoms[gb]
            if a["priority"] != b["priority"]:
                w = a if a["priority"] > b["priority"] else b
                fr, sr, si, cd = w["friction"], w["solref"], w["solimp"], w["condim"]
            else:
                mix = a["solmix"] / (a["solmix"] + b["solmix"]) if (a["solmix"] + b["solmix"]) > MJ_MINVAL else 0.5
                fr = np.maximum(a["friction"], b["friction"])
                sr = mix * a["solref"] + (1 - mix) * b["solref"] if (a["solref"][0] > 0 and b["solref"][0] > 0) \
                    else np.minimum(a["solref"], b["solref"])
                si = mix * a["solimp"] + (1 - mix) * b["solimp"]
                cd = max(a["condim"], b["condim"])
            if cd != 3:
                raise NotImplementedError("only condim=3 contacts are supported")
            pair_fn.append(fn)
            pair_g1.append(ga)
            pair_g2.append(gb)
            pair_conadr.append(ncon)
            pair_mu.append(fr[0])
            pair_solref.append(sr)
            pair_solimp.append(si)
            pair_margin.append(max(a["margin"], b["margin"]) - max(a["gap"], b["gap"]))
            ncon += _PAIR_NCON[fn]
    npair = len(pair_fn)
    A.update(
        pair_fn=np.array(pair_fn, dtype=i32), pair_geom1=np.array(pair_g1, dtype=i32),
        pair_geom2=np.array(pair_g2, dtype=i32), pair_conadr=np.array(pair_conadr, dtype=i32),
        pair_friction=np.array(pair_mu, dtype=np.float64),
        pair_solref=np.array(pair_solref, dtype=np.float64).reshape(npair, 2),
        pair_solimp=np.array(pair_solimp, dtype=np.float64).reshape(npair, 5),
        pair_includemargin=np.array(pair_margin, dtype=np.float64),
    )
    # joint-limit rows, in joint order (MJX _instantiate_limit_slide_hinge)
    limit_jnt = np.array([j for j in range(njnt) if jnts[j]["limited"] and jnts[j]["type"] == JNT_HINGE], dtype=i32)
    A.update(limit_jntid=limit_jnt)
    nlimit = len(limit_jnt)

    m.nq, m.nv, m.nu, m.na, m.nbody, m.njnt, m.ngeom = nq, nv, nu, na, nbody, njnt, ngeom
    m.nM, m.npair, m.ncon, m.nlimit, m.nefc = nM, npair, ncon, nlimit, nlimit + 4 * ncon
    m.arrays = A
    m.names = dict(body=[b.name for b in bodies], joint=[j["name"] for j in jnts],
                   geom=[g["name"] for g in geoms], actuator=[a["name"] for a in acts])
    _set_const(m)
    return m


# --------------------------------------------------------------------------- mj_setConst
def kinematics_np(m: FlatModel, qpos: np.ndarray):
    """float64 forward kinematics + subtree COM at `qpos` (used by set_const and by tests)."""
    nb = m.nbody
    xpos = np.zeros((nb, 3))
    xquat = np.tile(np.array([1.0, 0, 0, 0]), (nb, 1))
    xanchor = np.zeros((m.njnt, 3))
    xaxis = np.zeros((m.njnt, 3))
    for b in range(1, nb):
        p = m.body_parentid[b]
        pos = xpos[p] + quat_to_mat(xquat[p]) @ m.body_pos[b]
        quat = quat_mul(xquat[p], m.body_quat[b])
        for j in range(m.body_jntadr[b], m.body_jntadr[b] + m.body_jntnum[b]):
            a = m.jnt_qposadr[j]
            if m.jnt_type[j] == JNT_FREE:
                pos = qpos[a:a + 3].copy()
                quat = qpos[a + 3:a + 7] / np.linalg.norm(qpos[a + 3:a + 7])
                xanchor[j] = pos
                xaxis[j] = [0, 0, 1]
            else:
                R = quat_to_mat(quat)
                xanchor[j] = R @ m.jnt_pos[j] + pos
                xaxis[j] = R @ m.jnt_axis[j]
                quat = quat_mul(quat, axis_angle_quat(m.jnt_axis[j], qpos[a] - m.qpos0[a]))
                pos = xanchor[j] - quat_to_mat(quat) @ m.jnt_pos[j]
        xpos[b], xquat[b] = pos, quat / np.linalg.norm(quat)
    xipos = np.stack([xpos[b] + quat_to_mat(xquat[b]) @ m.body_ipos[b] for b in range(nb)])
    ximat = np.stack([quat_to_mat(quat_mul(xquat[b], m.body_iquat[b])) for b in range(nb)])
    return xpos, xquat, xipos, ximat, xanchor, xaxis


def mass_matrix_np(m: FlatModel, qpos: np.ndarray):
    """Dense float64 joint-space inertia via M = sum_b J_b^T I_b J_b (independent of the CRB formulation
    used by the oracle / kernels, so it doubles as a cross-check)."""
    xpos, xquat, xipos, ximat, xanchor, xaxis = kinematics_np(m, qpos)
    nv = m.nv
    M = np.zeros((nv, nv))
    jacs = []
    for b in range(m.nbody):
        jp_, jr_ = jac_point_np(m, xquat, xanchor, xaxis, b, xipos[b])
        jacs.append((jp_, jr_))
        if b == 0 or m.body_mass[b] <= 0:
            continue
        Iw = ximat[b] @ np.diag(m.body_inertia[b]) @ ximat[b].T
        M += m.body_mass[b] * jp_.T @ jp_ + jr_.T @ Iw @ jr_
    M += np.diag(m.dof_armature)
    return M, jacs


def jac_point_np(m: FlatModel, xquat, xanchor, xaxis, body: int, point: np.ndarray):
    """3 x nv translational and rotational Jacobians of `point` attached to `body` (world frame)."""
    nv = m.nv
    jacp = np.zeros((3, nv))
    jacr = np.zeros((3, nv))
    d = m.body_lastdof[body]
    while d >= 0:
        j = m.dof_jntid[d]
        if m.jnt_type[j] == JNT_FREE:
            k = d - m.jnt_dofadr[j]
            if k < 3:
                jacp[k, d] = 1.0
            else:
                ax = quat_to_mat(xquat[m.jnt_bodyid[j]])[:, k - 3]
                jacr[:, d] = ax
                jacp[:, d] = np.cross(ax, point - xanchor[j])
        else:
            jacr[:, d] = xaxis[j]
            jacp[:, d] = np.cross(xaxis[j], point - xanchor[j])
        d = m.dof_parentid[d]
    return jacp, jacr


def _set_const(m: FlatModel):
    """mj_setConst at qpos0: dof_invweight0, body_invweight0, stat.meaninertia."""
    nv = m.nv
    M, jacs = mass_matrix_np(m, m.qpos0)
    Minv = np.linalg.inv(M) if nv else np.zeros((0, 0))
    m.meaninertia = float(np.trace(M) / nv) if nv else 1.0
    body_invweight0 = np.zeros((m.nbody, 2))
    body_weld_static = np.array([m.body_lastdof[b] < 0 for b in range(m.nbody)])
    for b in range(1, m.nbody):
        if body_weld_static[b]:
            continue
        jp_, jr_ = jacs[b]
        At = jp_ @ Minv @ jp_.T
        Ar = jr_ @ Minv @ jr_.T
        body_invweight0[b] = [np.trace(At) / 3.0, np.trace(Ar) / 3.0]
    dof_invweight0 = np.zeros(nv)
    for j in range(m.njnt):
        a = m.jnt_dofadr[j]
        if m.jnt_type[j] == JNT_FREE:
            dof_invweight0[a:a + 3] = np.mean(np.diag(Minv)[a:a + 3])
            dof_invweight0[a + 3:a + 6] = np.mean(np.diag(Minv)[a + 3:a + 6])
        else:
            dof_invweight0[a] = Minv[a, a]
    m.arrays["body_invweight0"] = body_invweight0
    m.arrays["dof_invweight0"] = dof_invweight0
