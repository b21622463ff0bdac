"""Indirect shooting: oracle self-checks, the kernel body in the host simulation vs the oracle, the IPM's
init-constraint duals vs the oracle's, and (gpu) the CUDA kernel through the C ABI."""
import numpy as np
import pytest

from util import gb, orc, to_oracle, hostsim_iterate_with_duals, hostsim_shoot
from gusto_oracle import shooting as sh
from gusto_oracle.scp import solve_subproblem, cost_true, convergence_metric

CASES = [("dubins", dict(B=3, N=30)), ("astrobeeSE3manifold", dict(B=2, N=60))]


def later_duals_oracle(bp, b, iters=6):
    """SCPS.dual after a few GuSTO iterations of the oracle: the costate guess the reference's loop hands to shooting."""
    p = to_oracle(bp, b)
    return p, orc.solve_gusto(p, max_iter=iters, force=True).dual


def first_duals_oracle(bp, b):
    p = to_oracle(bp, b)
    sp = bp.model.scp_params
    X0, U0 = bp.init_traj_straightline()
    toggle = sp[0] / 8 + bp.model.clearance
    r = solve_subproblem(p, X0[b], U0[b], sp[1], sp[0], toggle, sp[3])[-1]
    n = p.model.n_x
    return p, np.array(r.nu[(p.N - 1) * n:p.N * n])


@pytest.mark.parametrize("name", ["dubins", "astrobeeSE3manifold"])
def test_oracle_costate_equations_are_the_hamiltonian_gradient(name):
    """d p/dt = -(df/dx)' p for the state equations the reference integrates (with u = get_control(x, p) held fixed the
    gyroscopic term of the manifold model is the one the reference comments out: J is isotropic, so it vanishes)."""
    m = orc.get_model(name)
    rng = np.random.default_rng(3)
    n = m.n_x
    y = rng.normal(size=2 * n)
    if name == "astrobeeSE3manifold":
        y[6:10] /= np.linalg.norm(y[6:10])
    d = sh.shooting_ode(m, y)
    u = sh.get_control(m, y[:n], y[n:])
    np.testing.assert_allclose(d[:n], orc.f_dyn(m, y[None, :n], u[None])[0], rtol=0, atol=1e-14)
    A = orc.A_dyn(m, y[None, :n])[0]
    np.testing.assert_allclose(d[n:], -A.T @ y[n:], rtol=0, atol=1e-13)


def test_oracle_rk4_matches_scipy_on_dubins():
    from scipy.integrate import solve_ivp
    m = orc.get_model("dubins")
    x0 = np.array([2.0, 2.0, 2.0]); p0 = np.array([0.2, -0.5, 2.5])
    ref = solve_ivp(lambda t, y: sh.shooting_ode(m, y), (0, 10.0), np.concatenate([x0, p0]), rtol=1e-11, atol=1e-12).y[:, -1]
    got = sh.integrate(m, x0, p0, 10.0, 30, 8)[0]
    assert np.max(np.abs(got - ref)) < 1e-5


@pytest.mark.parametrize("name,kw", CASES)
def test_ipm_body_exports_the_init_duals(name, kw):
    bp = gb.problems.CONFIGS[name](**kw)
    sp = bp.model.scp_params
    X0, U0 = bp.init_traj_straightline()
    hs = hostsim_iterate_with_duals(bp, X0, U0, sp[1], sp[0], stages=3)
    for b in range(bp.B):
        _, nu0 = first_duals_oracle(bp, b)
        assert hs["info"][b, 0] == 0
        # manifold: the multipliers of the (regularised, rank-deficient) quaternion rows are only determined to ~1e-5
        tol = 1e-6 if name == "dubins" else 2e-4
        assert np.max(np.abs(hs["dual"][b] - nu0)) <= tol * max(1.0, np.max(np.abs(nu0))), (hs["dual"][b], nu0)


@pytest.mark.parametrize("name,kw", CASES)
def test_shooting_body_matches_oracle(name, kw):
    bp = gb.problems.CONFIGS[name](**kw)
    X0, _ = bp.init_traj_straightline()
    xg = 0.5 * (bp.goal_lo + bp.goal_hi)
    duals = later_duals_oracle if name == "dubins" else first_duals_oracle
    p0 = np.stack([duals(bp, b)[1] for b in range(bp.B)])
    out, Xs, Us, Ps = hostsim_shoot(bp, p0, xg, X0)
    nchk = 0
    for b in range(bp.B):
        p = to_oracle(bp, b)
        r = sh.solve_shooting(p.model, p.x_init, xg[b], p0[b], p.tf, p.N)
        assert (r["status"] == "Optimal") == (out[b, 0] == 0)
        if r["status"] != "Optimal":
            assert np.isnan(out[b, 3]) and np.isnan(out[b, 4]) and np.array_equal(Xs[b], X0[b])
            continue
        if r["iters"] > 20:            # a long damped path is not reproducible to the last bit; the solution still has to be one
            assert np.max(np.abs(Xs[b][-1] - xg[b])) <= 1e-3
            continue
        nchk += 1
        assert out[b, 1] == r["iters"]
        assert out[b, 2] <= 1e-3 and abs(out[b, 2] - r["fnorm"]) < 1e-6
        assert np.max(np.abs(Xs[b] - r["X"])) < 1e-6 and np.max(np.abs(Us[b] - r["U"])) < 1e-6
        # the quaternion costate has a gauge direction (|q| = 1 is invariant: dx(tf)/dp0 is singular along it) that neither
        # the state nor the control sees, so it is not comparable; every other costate is
        keep = [i for i in range(p.model.n_x) if not (name == "astrobeeSE3manifold" and 6 <= i <= 9)]
        assert np.max(np.abs(Ps[b][:, keep] - r["P"][:, keep])) < 1e-6
        assert abs(out[b, 3] - cost_true(p, r["U"])) <= 1e-6 * max(1.0, out[b, 3])
        assert abs(out[b, 4] - convergence_metric(r["X"], X0[b])) < 1e-7
        assert np.max(np.abs(Xs[b][-1] - xg[b])) <= 1e-3             # the boundary-value problem is solved
    assert nchk >= 1


def test_shooting_reports_divergence():
    """An unreachable goal within the iteration cap is :Diverged, SS.traj untouched, NaN histories (shooting.jl:41-47)."""
    bp = gb.problems.CONFIGS["dubins"](B=1, N=30)
    X0, _ = bp.init_traj_straightline()
    xg = np.array([[500.0, 500.0, 0.0]])                              # v * tf = 20: cannot be reached
    out, Xs, _, _ = hostsim_shoot(bp, np.zeros((1, 3)), xg, X0, max_iter=5)
    assert out[0, 0] == 1 and np.isnan(out[0, 3]) and np.isnan(out[0, 4]) and np.array_equal(Xs, X0)
    p = to_oracle(bp, 0)
    assert sh.solve_shooting(p.model, p.x_init, xg[0], np.zeros(3), p.tf, p.N, max_iter=5)["status"] == "Diverged"


# ---------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name,kw", [("dubins", dict(B=16, N=30)), ("astrobeeSE3manifold", dict(B=8, N=60))])
def test_gpu_shooting_matches_oracle(name, kw, host):
    bp = gb.problems.CONFIGS[name](**kw)
    eng = host.Engine(bp)
    X0, U0 = bp.init_traj_straightline()
    eng.set_trajectory(X0, U0)
    for _ in range(6 if name == "dubins" else 1):                     # dubins: a few forced GuSTO iterations first
        _, info = eng.iterate()
        eng.accept(np.ones(bp.B, np.uint8))
    duals = eng.get_duals()
    Xacc, _ = eng.get_trajectory()                                    # SS.traj starts as the context's trajectory
    xg = 0.5 * (bp.goal_lo + bp.goal_hi)
    out = eng.shoot(None, xg)                                         # p0 = duals of the last solve, on the device
    Xs, Us, Ps = eng.get_shooting_trajectory()
    nchk = 0
    for b in range(min(bp.B, 4)):
        p = to_oracle(bp, b)
        assert info[b, 0] == 0
        if name != "dubins":
            nu0 = first_duals_oracle(bp, b)[1]
            assert np.max(np.abs(duals[b] - nu0)) <= 2e-4 * max(1.0, np.max(np.abs(nu0)))
        r = sh.solve_shooting(p.model, p.x_init, xg[b], duals[b], p.tf, p.N)
        assert (r["status"] == "Optimal") == (out[b, 0] == 0)
        if r["status"] == "Optimal" and r["iters"] <= 20:
            assert out[b, 1] == r["iters"]
            assert np.max(np.abs(Xs[b] - r["X"])) < 1e-6 and np.max(np.abs(Us[b] - r["U"])) < 1e-6
            assert abs(out[b, 3] - cost_true(p, r["U"])) <= 1e-6 * max(1.0, out[b, 3])
            assert abs(out[b, 4] - convergence_metric(r["X"], Xacc[b])) < 1e-7
            nchk += 1
    assert nchk >= 1
    ok = out[:, 0] == 0
    assert np.all(np.max(np.abs(Xs[ok][:, -1] - xg[ok]), axis=-1) <= 1e-3)     # every converged instance hits its goal
    # explicit p0 from the host gives the same answer as the device-resident duals
    out2 = eng.shoot(duals, xg)
    assert np.array_equal(out2[:, :2], out[:, :2])
    eng.close()


@pytest.mark.gpu
def test_gpu_scp_shooting_loop_dubins(host):
    """solve_SCPshooting! batched: shooting converges for instances whose SCP is still iterating, and TOS.traj is then
    the shooting trajectory (reaches the goal, cheaper than the SCP iterate it replaces)."""
    bp = gb.problems.CONFIGS["dubins"](B=32, N=30)
    eng = host.Engine(bp)
    SS = host.solve_scp_shooting_batch(eng, max_iter=30)
    assert SS.attempts >= 2 and SS.converged.sum() >= 1
    xg = 0.5 * (bp.goal_lo + bp.goal_hi)
    assert np.all(np.max(np.abs(SS.X[SS.converged][:, -1] - xg[SS.converged]), axis=-1) <= 1e-3)
    assert np.all(SS.SCPS.iterations[SS.converged] <= 30)
    with pytest.raises(host.GustoError):
        e2 = host.Engine(gb.problems.CONFIGS["astrobeeSE3"](B=1, N=20)); e2.shoot()
    eng.close()
