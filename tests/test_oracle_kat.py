"""The CPU oracle and the MJCF loader against everything the reference pins for this path (SURVEY.md section 4 /
8c): notebook-derived known answers (shapes, contact ordering, actuator-bias vector, unit root quaternion) and
internal invariants.  PARITY UNPINNED beyond these: the reference ships no tests and mujoco / mjx cannot be
installed here."""
import json
import os

import numpy as np
import pytest

from conftest import ROOT, load_asset

KAT = json.load(open(os.path.join(ROOT, "tests", "golden", "notebook_kat.json")))


def test_shape_kat(rodent0):
    """torchrl_explore.ipynb:63-152 prints every field shape of the MJX state for this topology."""
    m = rodent0
    assert (m.nq, m.nv, m.nu, m.na, m.nbody, m.njnt, m.ngeom) == (74, 73, 30, 30, 66, 68, 101)
    assert (m.ncon, m.nefc, m.nlimit) == (34, 203, 67)
    assert m.nM == 1119  # lower-triangular nnz of the tree-sparse mass matrix


@pytest.mark.parametrize("name,dims", [
    ("rodent_new", dict(nbody=67, nq=74, nv=73, nu=30, ncon=57, nefc=295)),
    ("rodent_optimized", dict(nbody=66, nq=74, nv=73, nu=30, ncon=59, nefc=303)),
    ("rodent_pair", dict(nbody=133, nq=148, nv=146, nu=60, ncon=114, nefc=590)),
])
def test_model_family_dims(name, dims):
    """SURVEY Appendix A (derived from the MJCF files with default-class resolution)."""
    m = load_asset(name)
    for k, v in dims.items():
        assert getattr(m, k) == v, (name, k)


def test_contact_order_kat(rodent0):
    """torchrl_explore.ipynb:658-661: contact.link_idx of geom2 -- capsule pairs first (2 contacts each, geom order),
    then the sphere (hand) pairs."""
    m = rodent0
    expected = [10, 10] + [11] * 6 + [14, 14] + [15] * 6 + [24, 24, 35, 35] + [59] * 6 + [64] * 6 + [58, 63]
    link = []
    for p in range(m.npair):
        n = 2 if m.pair_fn[p] == 1 else 1
        link += [int(m.geom_bodyid[m.pair_geom2[p]]) - 1] * n
        assert m.geom_bodyid[m.pair_geom1[p]] == 0  # geom1 is the floor (world, link -1)
    assert link == expected


def test_actuator_bias_kat(rodent0, oracle_mod):
    """Env_step.ipynb:295-313 (qfrc_actuator) with data.q (:1612-1630) at reset(PRNGKey(0)): act = 0 so
    qfrc_actuator = biasprm0 + biasprm1 * qpos_j on the actuated dofs.  Pins joint / dof / qpos addressing."""
    from brax_rodent_run_b200 import model_blob
    o = oracle_mod.Oracle(model_blob.pack(rodent0), "f64")
    o.init(np.array(KAT["qpos"]), np.zeros(rodent0.nv))
    got, want = o.get("qfrc_actuator"), np.array(KAT["qfrc_actuator"])
    assert np.count_nonzero(want) == 30
    assert np.abs(got - want).max() < 1e-7


def test_unit_quaternion_kat(rodent0, oracle_mod):
    """Env_step.ipynb:1612-1630: data.q[3:7] has unit norm after reset although +-0.01 noise was added: kinematics
    writes the normalised free-joint quaternion back into qpos."""
    q = np.array(KAT["qpos"])
    assert abs(np.linalg.norm(q[3:7]) - 1) < 1e-6
    from brax_rodent_run_b200 import model_blob
    o = oracle_mod.Oracle(model_blob.pack(rodent0), "f64")
    qpos = rodent0.qpos0 + 0.01
    o.init(qpos, np.zeros(rodent0.nv))
    assert abs(np.linalg.norm(o.get("qpos")[3:7]) - 1) < 1e-12
    assert np.allclose(o.get("qpos")[7:], qpos[7:])


def _state(m, seed):
    rng = np.random.default_rng(seed)
    qpos = m.qpos0 + rng.uniform(-.05, .05, m.nq)
    qpos[2] = 0.03
    return qpos, rng.uniform(-1, 1, m.nv)


def test_oracle_invariants(rodent0, oracle_mod):
    from brax_rodent_run_b200 import mjcf, model_blob
    m = rodent0
    o = oracle_mod.Oracle(model_blob.pack(m), "f64")
    o.set_options(0, 200, 50)
    qpos, qvel = _state(m, 0)
    o.init(qpos, qvel)
    M = o.get("qM").reshape(m.nv, m.nv)
    assert np.allclose(M, M.T) and np.linalg.eigvalsh(M).min() > 0
    # CRB mass matrix == sum_b J_b' I_b J_b (independent formulation in the loader)
    M2, _ = mjcf.mass_matrix_np(m, o.get("qpos"))
    assert np.abs(M - M2).max() / np.abs(M).max() < 1e-10
    assert np.allclose(M @ o.get("qacc_smooth"), o.get("qfrc_smooth"), rtol=1e-8, atol=1e-12)
    # (nearly) converged solver, 200 CG iterations: KKT residual M qacc - qfrc_smooth - J' f = 0, forces non-negative (pyramidal / limit rows)
    J = o.get("efc_J").reshape(m.nefc, m.nv)
    f = o.get("efc_force")
    res = M @ o.get("qacc") - o.get("qfrc_smooth") - J.T @ f
    assert np.abs(res).max() < 1e-4 * max(1.0, np.abs(o.get("qfrc_smooth")).max())  # CG stops on its own tolerance
    assert f.min() >= 0
    # setConst: meaninertia = trace(M(qpos0)) / nv
    o0 = oracle_mod.Oracle(model_blob.pack(m), "f64")
    o0.init(m.qpos0, np.zeros(m.nv))
    assert abs(np.trace(o0.get("qM").reshape(m.nv, m.nv)) / m.nv - m.meaninertia) < 1e-12


def test_solver_cost_monotone(rodent0, oracle_mod):
    """More CG iterations never increase the constraint cost (exact line search on a convex objective)."""
    from brax_rodent_run_b200 import model_blob
    m = rodent0
    qpos, qvel = _state(m, 1)
    costs = []
    for it in (1, 2, 4, 8, 16):
        o = oracle_mod.Oracle(model_blob.pack(m), "f64")
        o.set_options(0, it, 50)
        o.init(qpos, qvel)
        costs.append(o.scalar("solver_cost"))
    assert all(b <= a + 1e-9 * abs(a) for a, b in zip(costs, costs[1:]))


def test_free_fall(rodent0, oracle_mod):
    """Contact-free drop with zero joint velocities: the root accelerates at g for the first substep."""
    from brax_rodent_run_b200 import model_blob
    m = rodent0
    o = oracle_mod.Oracle(model_blob.pack(m), "f64")
    qpos = m.qpos0.copy()
    qpos[2] = 1.0
    o.init(qpos, np.zeros(m.nv))
    assert (o.get("contact_dist") > 0).all()
    com0 = o.get("subtree_com")[3:6].copy()
    o.step(1)
    o.forward()
    com1 = o.get("subtree_com")[3:6]
    dt = m.timestep
    # semi-implicit Euler: after one step the COM has moved by about g dt^2 (velocity updated first).  Not exact: joint
    # armature (reflected rotor inertia) and the limit rows active at qpos0 redistribute the first-step acceleration.
    dz = com1[2] - com0[2]
    assert dz < 0 and abs(dz / (-9.81 * dt * dt) - 1) < 0.15
    assert abs(com1[0] - com0[0]) < 1e-5 and abs(com1[1] - com0[1]) < 1e-5


def test_oracle_golden_trajectory(rodent0, oracle_mod):
    """Regression fixture for the oracle itself (tests/golden/oracle_traj.npz, made by tools/make_oracle_golden.py)."""
    from brax_rodent_run_b200 import model_blob
    g = np.load(os.path.join(ROOT, "tests", "golden", "oracle_traj.npz"))
    o = oracle_mod.Oracle(model_blob.pack(rodent0), "f64")
    o.set_options(0, int(g["iterations"]), int(g["ls_iterations"]))
    o.init(g["qpos0"], g["qvel0"])
    for t in range(g["ctrl"].shape[0]):
        o.set("ctrl", g["ctrl"][t])
        o.step(1)
        assert np.abs(o.get("qpos") - g["qpos"][t]).max() < 1e-9
        assert np.abs(o.get("qvel") - g["qvel"][t]).max() < 1e-7
