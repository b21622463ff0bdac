"""Host-side outer-loop logic (accept/reject, Delta/omega schedule, convergence window) against the oracle's scalar
loop, and the multi-rank sharding / status all-gather over gloo (world_size 2, CPU)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from util import gb, orc, to_oracle, ROOT


def scalar_update(ev, solver_ok, Delta, omega, iterations, conv_prev, sp):
    """scp_gusto.jl:119-174 for ONE instance, written independently of host.gusto_update."""
    D0, w0, w_max, eps, rho0, rho1, b_succ, b_fail, g_fail, thr = sp
    conv, tr_ok, ineq_ok, rho = ev[0], ev[1] > 0.5, ev[2] > 0.5, ev[3]
    if not solver_ok:
        return dict(accept=False, Delta=Delta, omega=omega, iterations=iterations, done=True, converged=False, successful=False)
    if tr_ok:
        if rho > rho1:
            accept, Delta_n, omega_n = False, b_fail * Delta, omega
        else:
            accept = True
            Delta_n = min(b_succ * Delta, D0) if rho < rho0 else Delta
            omega_n = omega if ineq_ok else g_fail * omega
    else:
        accept, Delta_n, omega_n = False, Delta, g_fail * omega
    iterations += 1
    done = converged = successful = False
    if omega_n > w_max:
        done = True
    elif accept and iterations > 2 and conv + conv_prev <= thr:
        converged, successful, done = True, bool(ineq_ok), True
    return dict(accept=accept, Delta=Delta_n, omega=omega_n, iterations=iterations, done=done, converged=converged, successful=successful)


def test_vectorised_update_equals_scalar_reference_logic(host):
    rng = np.random.default_rng(0)
    sp = gb.models.AstrobeeSE3().scp_params
    B = 4000
    ev = np.zeros((B, 8))
    ev[:, 0] = rng.choice([1e-4, 4e-3, 0.02, 0.3], B); ev[:, 1] = rng.integers(0, 2, B); ev[:, 2] = rng.integers(0, 2, B)
    ev[:, 3] = rng.choice([1e-3, 0.02, 0.2], B)
    ok = rng.random(B) > 0.05
    active = rng.random(B) > 0.1
    Delta = rng.choice([10.0, 5.0, 0.3], B); omega = rng.choice([1.0, 625.0, 5e9, 1e10], B)
    its = rng.integers(0, 6, B); cprev = rng.choice([1e-4, 4e-3, 0.02], B)
    st = host.gusto_update(ev, ok, active, Delta, omega, its, cprev, sp)
    for b in range(B):
        if not active[b]:
            assert not st["accept"][b] and st["Delta"][b] == Delta[b] and st["omega"][b] == omega[b] and st["iterations"][b] == its[b]
            continue
        r = scalar_update(ev[b], ok[b], Delta[b], omega[b], its[b], cprev[b], sp)
        assert bool(st["accept"][b]) == r["accept"] and st["Delta"][b] == r["Delta"] and st["omega"][b] == r["omega"]
        assert st["iterations"][b] == r["iterations"] and bool(st["done"][b]) == r["done"]
        assert bool(st["converged_now"][b]) == r["converged"] and bool(st["successful_now"][b]) == r["successful"]


@pytest.mark.parametrize("force", [False, True])
def test_device_update_kernel_logic_equals_host_update(host, force):
    """scp.cuh::scp_update_instance (the decision table of the device-resident loop, compiled for the host) against
    host.gusto_update on random evaluation records: accept, Delta, omega, iterations, done, converged, successful, statuses."""
    import ctypes
    from util import hostsim_lib
    rng = np.random.default_rng(3)
    sp = np.ascontiguousarray(gb.models.AstrobeeSE3().scp_params, dtype=np.float64)
    B = 5000
    ev = np.zeros((B, 8)); info = np.zeros((B, 8))
    ev[:, 0] = rng.choice([1e-4, 4e-3, 0.02, 0.3], B); ev[:, 1] = rng.integers(0, 2, B); ev[:, 2] = rng.integers(0, 2, B)
    ev[:, 3] = rng.choice([1e-3, 0.02, 0.2], B); ev[:, 4] = rng.random(B)
    info[:, 0] = rng.choice([0, 0, 0, 0, 1, 2, 3], B); info[:, 1] = rng.integers(5, 20, B); info[:, 4] = rng.random(B)
    active = rng.random(B) > 0.1
    Delta = rng.choice([10.0, 5.0, 0.3], B); omega = rng.choice([1.0, 625.0, 5e9, 1e10], B)
    its = rng.integers(0, 6, B).astype(np.int64); cprev = rng.choice([1e-4, 4e-3, 0.02], B)
    st = host.gusto_update(ev, host.solver_status_ok(info[:, 0]), active, Delta, omega, its, cprev, sp, force)
    d2, w2, it2, c2 = Delta.copy(), omega.copy(), its.astype(np.int32), cprev.copy()
    jt = rng.random(B); jf = rng.random(B); jt0 = jt.copy()
    cv = np.zeros(B, np.uint8); su = np.zeros(B, np.uint8); rec = np.zeros((B, host.HIST_W)); fl = np.zeros(B, np.int32)
    dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    ip = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
    bp_ = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))
    act8 = active.astype(np.uint8)
    assert hostsim_lib().hostsim_scp_update(B, dp(ev), dp(info), dp(sp), int(force), bp_(act8), dp(d2), dp(w2), ip(it2), dp(c2), dp(jt), dp(jf),
                                            bp_(cv), bp_(su), dp(rec), ip(fl)) == 0
    assert np.array_equal((fl & 1) > 0, st["accept"]) and np.array_equal(d2, st["Delta"]) and np.array_equal(w2, st["omega"])
    assert np.array_equal(it2, st["iterations"]) and np.array_equal(cv > 0, st["converged_now"]) and np.array_equal(su > 0, st["successful_now"])
    assert np.array_equal(((fl & 2) > 0)[active], st["done"][active]) and np.all((fl & 2)[~active] > 0)
    assert np.array_equal(rec[:, host.H_SCP_STATUS].astype(np.int32), st["status"])
    assert np.array_equal(jt, np.where(st["accept"], ev[:, 4], jt0))
    assert np.array_equal(c2, np.where(st["run"], ev[:, 0], cprev))


def test_shard_is_a_contiguous_partition(pkg):
    bp = pkg.problems.config_astrobee_se3(B=37, N=10, seed=2)
    parts = [bp.shard(r, 4) for r in range(4)]
    assert sum(p.B for p in parts) == 37
    assert np.array_equal(np.concatenate([p.x_init for p in parts]), bp.x_init)
    assert np.array_equal(np.concatenate([p.goal_lo for p in parts]), bp.goal_lo)


GLOO_WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
from util import gb, hostsim_iterate
host = gb.engine()
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
bp_all = gb.problems.config_freeflyer_se2(B=6, N=12, seed=4)
bp = bp_all.shard(rank, world)
sp = bp.model.scp_params
X, U = bp.init_traj_straightline()
B = bp.B
Delta, omega = np.full(B, sp[0]), np.full(B, sp[1])
its, cprev, active = np.zeros(B, np.int64), np.zeros(B), np.ones(B, bool)
flags_all = torch.zeros(bp_all.B, dtype=torch.uint8)
for it in range(12):
    hs = hostsim_iterate(bp, X, U, omega, Delta)          # kernel bodies on the host (test harness only)
    st = host.gusto_update(hs["eval"], hs["info"][:, 0] == 0, active, Delta, omega, its, cprev, sp)
    acc = st["accept"]
    X[acc], U[acc] = hs["Xn"][acc], hs["Un"][acc]
    Delta, omega, its, cprev = st["Delta"], st["omega"], st["iterations"], np.where(st["run"], hs["eval"][:, 0], cprev)
    active = active & ~st["done"]
    # the path's only collective: all-gather of the per-instance status bytes
    mine = torch.from_numpy((~active).astype(np.uint8))
    parts = [torch.zeros(bp_all.shard(r, world).B, dtype=torch.uint8) for r in range(world)]
    dist.all_gather(parts, mine)
    if bool(torch.cat(parts).all()):
        break
out = np.concatenate([X.reshape(B, -1), its[:, None].astype(float)], axis=1)
np.save(os.path.join(sys.argv[2], f"rank{rank}_of{world}.npy"), out)
dist.destroy_process_group()
'''


@pytest.mark.slow
def test_two_rank_gloo_run_equals_single_rank_run(tmp_path):
    """Sharding the batch over 2 ranks (status all-gather per outer iteration) gives bit-identical per-instance results."""
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER)
    outs = {}
    for world in (1, 2):
        d = tmp_path / f"w{world}"
        d.mkdir()
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
               "--master-port", str(29511 + world), str(script), ROOT, str(d)]
        subprocess.check_call(cmd, env=dict(os.environ, OMP_NUM_THREADS="1"), timeout=600)
        outs[world] = np.concatenate([np.load(d / f"rank{r}_of{world}.npy") for r in range(world)])
    assert np.array_equal(outs[1], outs[2])


def test_box_goals_reach_the_solver_unchanged(host):
    """Round 1 presolved a narrow BoxGoal (astrobeeSE3manifold notebook, q +- 1e-4) into a PointGoal on the host; since the
    Riccati solve needs no such help the goal table is handed over exactly as the reference's GoalSet states it."""
    bp = gb.problems.config_astrobee_se3_manifold(B=3, N=12)
    cfg, _ = host.make_config(bp)
    assert list(cfg.goal_type[:13]) == list(bp.goal_type) and list(bp.goal_type[6:10]) == [2, 2, 2, 2]
    assert not hasattr(host, "presolve_goals")
    assert np.allclose(bp.goal_hi[:, 6:10] - bp.goal_lo[:, 6:10], 2e-4)


def test_almost_optimal_solver_status_continues_the_scp(host):
    """scp_gusto.jl:107: OPTIMAL / LOCALLY_SOLVED / ALMOST_LOCALLY_SOLVED go on, every other status returns."""
    st = np.array([0, 1, 2, 3], dtype=np.float64)
    assert list(host.solver_status_ok(st)) == [True, False, False, True]
