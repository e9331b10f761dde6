"""Golden vectors from the UNMODIFIED reference environment (Rodent_Env_Brax.py on mujoco-mjx / brax / jax).

This image has no mujoco / mjx / brax / jax, so the script cannot run here; it is committed so that anybody with the
reference's dependencies can produce the fixtures that tests/test_mjx_golden.py consumes when they are present:

    pip install mujoco mujoco-mjx brax jax dm_control           # the reference's (unpinned) dependencies
    python tools/make_mjx_golden.py --reference /path/to/Brax-Rodent-Run --model rodent_0 --iterations 8
    python tools/make_mjx_golden.py --reference /path/to/Brax-Rodent-Run --model rodent_new --iterations 6
    -> tests/golden/mjx_<model>_cg<it>.npz

What is driven (nothing in the reference tree is edited; only the module-level `_XML_PATH` of Rodent_Env_Brax.py:16 is
pointed at the model asked for, and the process runs with the reference checkout as its working directory because that path
is relative):
    Rodent(track_pos, solver="cg", iterations, ls_iterations)            Rodent_Env_Brax.py:21-69
    reset(PRNGKey(seed))                                                  :71-96   (pipeline_init = mjx.forward)
    step(state, action) with n_frames = 1 (one mjx.step) and n_frames = 10  :98-136
    a 100-step random-action trajectory with the env's own step
and, from mujoco.MjModel.from_xml_path (Rodent_Env_Brax.py:41), the compiled-model constants the in-repo MJCF loader must
reproduce (body_mass / body_inertia / body_ipos / body_iquat / body_invweight0 / dof_invweight0 / stat.meaninertia ...).

`--source oracle` writes a file of the SAME layout from the in-repo CPU oracle instead.  It exists so that the comparison
code of tests/test_mjx_golden.py is exercised in this image (self-check); it is not reference data and is never committed
under the mjx_ prefix.

File layout (npz, float64 / int32 arrays):
    meta_model, meta_iterations, meta_ls_iterations, meta_source, meta_versions, track_pos [T,3]
    const_<name>                            compiled-model constants (mjx source only)
    reset_cur_frame, reset_obs, reset_<f>   state + forward intermediates after reset           (<f> in FIELDS)
    sub1_action, sub1_<f>                   after ONE mjx.step from the reset state
    step10_obs, step10_reward, step10_done, step10_<f>   after one env.step (10 substeps) from the reset state, same action
    traj_actions [N,nu], traj_qpos [N,nq], traj_qvel [N,nv], traj_reward [N], traj_done [N], traj_obs10 [N/10,obs]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")

# mjx.Data / brax.mjx.State fields dumped when present (name in file = name here)
FIELDS = ("qpos", "qvel", "act", "qacc_warmstart", "ctrl", "xpos", "xquat", "xmat", "xipos", "subtree_com", "cinert", "cdof",
          "cvel", "qM", "qfrc_bias", "qfrc_passive", "qfrc_actuator", "qfrc_smooth", "qacc_smooth", "qfrc_constraint", "qacc",
          "efc_J", "efc_D", "efc_aref", "efc_force")
CONTACT_FIELDS = ("dist", "pos", "frame", "friction", "solref", "solimp", "geom1", "geom2")
MODEL_CONSTS = ("body_mass", "body_inertia", "body_ipos", "body_iquat", "body_pos", "body_quat", "body_invweight0",
                "body_subtreemass", "body_parentid", "body_jntadr", "body_jntnum", "dof_invweight0", "dof_armature", "dof_damping",
                "jnt_stiffness", "jnt_range", "jnt_solref", "jnt_solimp", "jnt_pos", "jnt_axis", "jnt_type", "qpos0",
                "qpos_spring", "geom_type", "geom_size", "geom_pos", "geom_quat", "geom_friction", "geom_solref",
                "geom_solimp", "geom_priority", "geom_condim", "geom_contype", "geom_conaffinity", "geom_bodyid",
                "actuator_gainprm", "actuator_biasprm", "actuator_dynprm", "actuator_gear", "actuator_trnid")


def synthetic_track(n=250):
    """SURVEY section 8(d) config 2 (the track every parity test of this repo uses)."""
    return np.stack([0.002 * np.arange(n), np.zeros(n), np.full(n, 0.055)], 1).astype(np.float32)


def dump_data(out, prefix, data):
    """mjx.Data -> out[prefix + field]"""
    for f in FIELDS:
        v = getattr(data, f, None)
        if v is not None:
            out[prefix + f] = np.asarray(v, np.float64)
    c = getattr(data, "contact", None)
    if c is not None:
        for f in CONTACT_FIELDS:
            v = getattr(c, f, None)
            if v is not None:
                out[prefix + "contact_" + f] = np.asarray(v, np.float64)


def from_mjx(a, out):
    ref = os.path.abspath(a.reference)
    os.chdir(ref)  # _XML_PATH is relative (Rodent_Env_Brax.py:16)
    sys.path.insert(0, ref)
    import jax
    import mujoco
    from jax import numpy as jp
    import Rodent_Env_Brax as R  # the unmodified reference module

    R._XML_PATH = f"./models/{a.model}.xml"
    track = synthetic_track()
    kw = dict(solver="cg", iterations=a.iterations, ls_iterations=a.ls_iterations)
    env10 = R.Rodent(track_pos=jp.asarray(track), **kw)
    env1 = R.Rodent(track_pos=jp.asarray(track), n_frames=1, **kw)
    import brax
    out["meta_versions"] = np.array(f"mujoco {mujoco.__version__} jax {jax.__version__} brax {getattr(brax, '__version__', '?')}")
    mjm = mujoco.MjModel.from_xml_path(R._XML_PATH)
    for name in MODEL_CONSTS:
        out["const_" + name] = np.asarray(getattr(mjm, name), np.float64)
    out["const_meaninertia"] = np.float64(mjm.stat.meaninertia)
    out["const_dims"] = np.array([mjm.nq, mjm.nv, mjm.nu, mjm.na, mjm.nbody, mjm.njnt, mjm.ngeom], np.int32)

    reset10, step10, step1 = jax.jit(env10.reset), jax.jit(env10.step), jax.jit(env1.step)
    s0 = reset10(jax.random.PRNGKey(a.seed))
    out["reset_cur_frame"] = np.int32(s0.info["cur_frame"])
    out["reset_obs"] = np.asarray(s0.obs, np.float64)
    dump_data(out, "reset_", s0.pipeline_state)

    rng = np.random.default_rng(a.seed)
    nu = int(mjm.nu)
    action = rng.uniform(-1, 1, nu).astype(np.float32)
    out["sub1_action"] = action.astype(np.float64)
    s1 = step1(s0, jp.asarray(action))
    dump_data(out, "sub1_", s1.pipeline_state)
    s10 = step10(s0, jp.asarray(action))
    dump_data(out, "step10_", s10.pipeline_state)
    out["step10_obs"], out["step10_reward"], out["step10_done"] = (np.asarray(s10.obs, np.float64), np.float64(s10.reward),
                                                                   np.float64(s10.done))
    acts = rng.uniform(-1, 1, (a.traj, nu)).astype(np.float32)
    qs, vs, rs, ds, obs10 = [], [], [], [], []
    s = s0
    for t in range(a.traj):
        s = step10(s, jp.asarray(acts[t]))
        qs.append(np.asarray(s.pipeline_state.qpos, np.float64)); vs.append(np.asarray(s.pipeline_state.qvel, np.float64))
        rs.append(float(s.reward)); ds.append(float(s.done))
        if t % 10 == 9:
            obs10.append(np.asarray(s.obs, np.float64))
    out.update(traj_actions=acts.astype(np.float64), traj_qpos=np.array(qs), traj_qvel=np.array(vs), traj_reward=np.array(rs),
               traj_done=np.array(ds), traj_obs10=np.array(obs10))
    os.chdir(ROOT)


class _OracleData:
    """mjx.Data-shaped view of the oracle (dense efc rows, contact.* names)"""

    def __init__(self, o, m):
        self._o, self._m = o, m

    def __getattr__(self, f):
        o, m = self._o, self._m
        shapes = {"xpos": (-1, 3), "xquat": (-1, 4), "xmat": (-1, 9), "xipos": (-1, 3), "subtree_com": (-1, 3), "cinert": (-1, 10),
                  "cdof": (-1, 6), "cvel": (-1, 6), "qM": (m.nv, m.nv), "efc_J": (-1, m.nv)}
        if f == "contact":
            class C:
                dist = o.get("contact_dist"); pos = o.get("contact_pos").reshape(-1, 3); frame = o.get("contact_frame").reshape(-1, 3, 3)
            return C
        try:
            v = o.get(f)
        except KeyError:
            return None
        return v.reshape(shapes[f]) if f in shapes else v


def from_oracle(a, out):
    """Same layout from the in-repo oracle (self-check of the comparison code; NOT reference data)."""
    sys.path.insert(0, ROOT)
    from brax_rodent_run_b200 import mjcf, model_blob
    from oracle import oracle
    oracle.build()
    m = mjcf.FlatModel.load(os.path.join(ROOT, "brax_rodent_run_b200", "assets", a.model + ".npz"))
    track = synthetic_track()
    blob = model_blob.pack(m)
    mk = lambda nf: oracle.OracleRodentEnv(blob, (m.nq, m.nv, m.nu, m.nbody), track, iterations=a.iterations,
                                           ls_iterations=a.ls_iterations, n_frames=nf, precision="f64")
    rng = np.random.default_rng(a.seed)
    sf = int(rng.integers(0, 100))
    q0 = m.qpos0.copy(); q0[:3] = track[sf]
    q0 = q0 + rng.uniform(-.01, .01, m.nq); v0 = rng.uniform(-.01, .01, m.nv)
    e1, e10 = mk(1), mk(10)
    out["meta_versions"] = np.array("in-repo oracle (self-check)")
    out["reset_cur_frame"] = np.int32(sf)
    out["reset_obs"] = e10.reset(sf, q0, v0)
    e1.reset(sf, q0, v0)
    dump_data(out, "reset_", _OracleData(e10.o, m))
    rng = np.random.default_rng(a.seed)
    action = rng.uniform(-1, 1, m.nu).astype(np.float32).astype(np.float64)
    out["sub1_action"] = action
    e1.step(action)
    dump_data(out, "sub1_", _OracleData(e1.o, m))
    ob, r, d, _ = e10.step(action)
    dump_data(out, "step10_", _OracleData(e10.o, m))
    out["step10_obs"], out["step10_reward"], out["step10_done"] = ob, np.float64(r), np.float64(d)
    acts = rng.uniform(-1, 1, (a.traj, m.nu)).astype(np.float32).astype(np.float64)
    e = mk(10); e.reset(sf, q0, v0)
    qs, vs, rs, ds, obs10 = [], [], [], [], []
    for t in range(a.traj):
        ob, r, d, _ = e.step(acts[t])
        qs.append(e.o.get("qpos")); vs.append(e.o.get("qvel")); rs.append(r); ds.append(d)
        if t % 10 == 9:
            obs10.append(ob)
    out.update(traj_actions=acts, traj_qpos=np.array(qs), traj_qvel=np.array(vs), traj_reward=np.array(rs), traj_done=np.array(ds),
               traj_obs10=np.array(obs10))


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference", help="checkout of talmolab/Brax-Rodent-Run")
    ap.add_argument("--model", default="rodent_0", help="models/<model>.xml of the reference")
    ap.add_argument("--iterations", type=int, default=8)
    ap.add_argument("--ls-iterations", type=int, default=None)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--traj", type=int, default=100)
    ap.add_argument("--source", default="mjx", choices=["mjx", "oracle"])
    ap.add_argument("--out", default=None)
    a = ap.parse_args(argv)
    if a.ls_iterations is None:
        a.ls_iterations = a.iterations
    out = dict(meta_model=np.array(a.model), meta_iterations=np.int32(a.iterations), meta_ls_iterations=np.int32(a.ls_iterations),
               meta_source=np.array(a.source), track_pos=synthetic_track().astype(np.float64))
    (from_mjx if a.source == "mjx" else from_oracle)(a, out)
    path = a.out or os.path.join(ROOT, "tests", "golden", f"{'mjx' if a.source == 'mjx' else 'selfcheck'}_{a.model}_cg{a.iterations}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)
    return path


if __name__ == "__main__":
    main()
