# GuSTOB200.jl -- Julia host side of the B200-native GuSTO hot path.
#
# Drop-in for the `solve_method!` slot of `solve_SCP!` (reference: src/traj_opt.jl:47-60,59):
#
#     include("../src/GuSTO.jl")                       # the unchanged reference package
#     include("GuSTOB200.jl"); using .GuSTOB200
#     solve_SCP!(TOS, TOP, solve_gusto_b200!, init_traj_straightline, "B200")
#
# `solve_gusto_b200!` has the signature of `solve_gusto_jump!` (src/scp/scp_gusto.jl:49) and fills the same
# SCPSolution / SCPParam_GuSTO histories (types.jl:150-173, scp_gusto.jl:4-24).  The outer trust-region update and
# convergence test (scp_gusto.jl:119-174) run here in Julia; linearization, the convex subproblem and the evaluation
# scalars run in libgusto_b200.so through `ccall` (include/gusto_b200.h).  `solve_SCP_batch!` drives B independent
# problems that share robot / model / environment through one context with the same Julia loop; `solve_SCP_batch_device!`
# hands the whole outer loop to the library (gusto_scp_run) and is the multi-GPU entry point (one process per GPU).
#
# NOTE: Julia is not installed in the build or GPU containers of this project, so this file has been reviewed against
# include/gusto_b200.h and gusto.jl_b200/host.py (which implements the identical loop and IS tested) but never executed.
module GuSTOB200

export solve_trajopt_b200!, solve_gusto_b200!, solve_SCP_batch!, solve_SCP_batch_device!, solve_shooting_b200!, comm_unique_id, GustoContext

const LIB = get(ENV, "GUSTO_B200_LIB", joinpath(@__DIR__, "..", "libgusto_b200.so"))

const GUSTO_DUBINS, GUSTO_FREEFLYER_SE2, GUSTO_ASTROBEE_SE3, GUSTO_ASTROBEE_SE3_MANIFOLD = Int32(0), Int32(1), Int32(2), Int32(3)
const GOAL_FREE, GOAL_POINT, GOAL_BOX = Int32(0), Int32(1), Int32(2)
const EVAL_NOUT, SOLVE_NINFO = 8, 8

# Mirror of `gusto_config` (include/gusto_b200.h).  NTuple fields give the C layout.
struct GustoConfig
  model_id::Int32
  N::Int32
  B::Int32
  n_obs::Int32
  robot_params::NTuple{16,Float64}
  scp_params::NTuple{10,Float64}
  goal_type::NTuple{16,Int32}
  device::Int32
  ipm_max_iter::Int32
  ipm_nref::Int32
  ipm_tol::Float64
  ipm_delta_p::Float64
  ipm_delta_d::Float64
end

mutable struct GustoContext
  ptr::Ptr{Cvoid}
  B::Int
  N::Int
  x_dim::Int
  u_dim::Int
end

function check(ctx, rc::Int32)
  if rc != 0
    msg = unsafe_string(ccall((:gusto_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx === nothing ? C_NULL : ctx.ptr))
    error("gusto_b200 call failed ($rc): $msg")
  end
end

# ------------------------------------------------------------------------------------------------ flattening
# model_id from the dynamics-model type name (src/dynamics/*.jl); dispatch on names keeps this file loadable
# without the reference package.
function model_id(model)
  n = string(nameof(typeof(model)))
  n == "DubinsCar" && return GUSTO_DUBINS
  n == "FreeflyerSE2" && return GUSTO_FREEFLYER_SE2
  n == "AstrobeeSE3" && return GUSTO_ASTROBEE_SE3
  n == "AstrobeeSE3Manifold" && return GUSTO_ASTROBEE_SE3_MANIFOLD
  error("GuSTOB200: unsupported dynamics model $n")
end

# robot_params[16]: 0 mass | 1-3 Jxx,Jyy,Jzz | 4 radius | 5 v_max | 6 a_max | 7 w_max | 8 alpha_max | 9 clearance |
# 10 dubins v | 11 dubins k | 12-14 dubins x_max | 15 dubins u_max      (robot/astrobee3D.jl:16-30, freeflyer.jl:29-50)
function robot_params(robot, model)
  p = zeros(16)
  p[10] = model.clearance
  rn = string(nameof(typeof(robot)))
  if rn == "Astrobee3D"
    p[1] = robot.mass; p[2] = robot.J[1,1]; p[3] = robot.J[2,2]; p[4] = robot.J[3,3]
    p[5] = robot.r; p[6] = robot.hard_limit_vel; p[7] = robot.hard_limit_accel
    p[8] = robot.hard_limit_ω; p[9] = robot.hard_limit_α
  elseif rn == "Freeflyer"
    p[1] = robot.mass_ff; p[2] = p[3] = p[4] = robot.J_ff
    p[5] = robot.r; p[6] = robot.hard_limit_vel; p[7] = robot.hard_limit_accel
    p[8] = robot.hard_limit_ω; p[9] = robot.hard_limit_α
  elseif rn == "Car"
    p[1] = 1.0; p[11] = model.v; p[12] = model.k
    p[13:15] = model.x_max; p[16] = model.u_max
  else
    error("GuSTOB200: unsupported robot $rn")
  end
  return p
end

# Collision components in the reference's order keepout_zones..., obstacle_set... (types.jl:19).
# HyperRectangle -> (0, origin, origin + widths); HyperSphere -> (1, center, (r, 0, 0)).
function obstacle_table(env, model)
  kinds = Int32[]; a = Float64[]; b = Float64[]
  model_id(model) == GUSTO_DUBINS && return kinds, a, b      # dubins registers no obstacle rows (dubins_car.jl:184-226)
  for zone in (env.keepout_zones..., env.obstacle_set...)
    if hasproperty(zone, :widths)
      lo = Float64.(collect(zone.origin)); hi = Float64.(collect(zone.origin .+ zone.widths))
      push!(kinds, 0); append!(a, min.(lo, hi)); append!(b, max.(lo, hi))
    else
      push!(kinds, 1); append!(a, Float64.(collect(zone.center))); append!(b, [Float64(zone.r), 0.0, 0.0])
    end
  end
  return kinds, a, b
end

# Final-time goals -> per-coordinate (type, lo, hi)   (goals.jl; registries e.g. astrobee_se3.jl:339-345)
function flatten_goals(goal_set, x_dim, tf_guess)
  gtype = zeros(Int32, 16); lo = zeros(x_dim); hi = zeros(x_dim)
  for (t, goal) in goal_set.goals
    t == tf_guess || continue
    ind = collect(goal.ind_coordinates)
    if string(nameof(typeof(goal.params))) == "PointGoal"
      gtype[ind] .= GOAL_POINT; lo[ind] = goal.params.point; hi[ind] = goal.params.point
    else
      gtype[ind] .= GOAL_BOX; lo[ind] = goal.params.lower_bound; hi[ind] = goal.params.upper_bound
    end
  end
  # BoxGoal rows go to the solver as they are (csbci_goal_constraints, dynamics.jl:37-42): no presolve since round 2
  return gtype, lo, hi
end

solver_status_symbol(s) = s == 0 ? :OPTIMAL : s == 1 ? :ITERATION_LIMIT : s == 3 ? :ALMOST_OPTIMAL : :NUMERICAL_ERROR

scp_params(alg, param) = Float64[alg.Δ0, alg.ω0, alg.ω_max, alg.ε, alg.ρ0, alg.ρ1, alg.β_succ, alg.β_fail, alg.γ_fail,
                                 param.convergence_threshold]

# ---------------------------------------------------------------------------------------------------- context
function GustoContext(robot, model, env, N::Int, B::Int, goal_type, alg, param; device::Int=0)
  kinds, a, b = obstacle_table(env, model)
  cfg = GustoConfig(model_id(model), N, B, length(kinds), Tuple(robot_params(robot, model)), Tuple(scp_params(alg, param)),
                    Tuple(goal_type), device, 0, 0, 0.0, 0.0, 0.0)
  out = Ref{Ptr{Cvoid}}(C_NULL)
  rc = GC.@preserve kinds a b ccall((:gusto_create, LIB), Int32,
      (Ref{GustoConfig}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ref{Ptr{Cvoid}}), cfg, kinds, a, b, out)
  rc == 0 || error("gusto_create failed ($rc): " * unsafe_string(ccall((:gusto_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
  ctx = GustoContext(out[], B, N, model.x_dim, model.u_dim)
  finalizer(c -> (c.ptr != C_NULL && ccall((:gusto_destroy, LIB), Int32, (Ptr{Cvoid},), c.ptr); c.ptr = C_NULL), ctx)
  return ctx
end

set_problems!(ctx, x_init, lo, hi, tf) = GC.@preserve x_init lo hi tf check(ctx, ccall((:gusto_set_problems, LIB), Int32,
    (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), ctx.ptr, x_init, lo, hi, tf))
set_trajectory!(ctx, X, U) = GC.@preserve X U check(ctx, ccall((:gusto_set_trajectory, LIB), Int32,
    (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), ctx.ptr, X, U))
set_candidate!(ctx, X, U) = GC.@preserve X U check(ctx, ccall((:gusto_set_candidate, LIB), Int32,
    (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), ctx.ptr, X, U))
get_trajectory!(ctx, X, U) = GC.@preserve X U check(ctx, ccall((:gusto_get_trajectory, LIB), Int32,
    (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), ctx.ptr, X, U))
set_penalties!(ctx, ω, Δ) = GC.@preserve ω Δ check(ctx, ccall((:gusto_set_penalties, LIB), Int32,
    (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), ctx.ptr, ω, Δ))
set_active!(ctx, act::Vector{UInt8}) = GC.@preserve act check(ctx, ccall((:gusto_set_active, LIB), Int32,
    (Ptr{Cvoid}, Ptr{UInt8}), ctx.ptr, act))
linearize!(ctx) = check(ctx, ccall((:gusto_linearize, LIB), Int32, (Ptr{Cvoid},), ctx.ptr))
evaluate!(ctx, out) = GC.@preserve out check(ctx, ccall((:gusto_evaluate, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), ctx.ptr, out))
iterate!(ctx, out, info) = GC.@preserve out info check(ctx, ccall((:gusto_iterate, LIB), Int32,
    (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), ctx.ptr, out, info))
# one iteration with host-resident trajectories: uploads (X, U, ω, Δ, active), runs K1 -> K3 -> K4, downloads (out, info, Xn, Un)
iterate_host!(ctx, X, U, ω, Δ, act::Vector{UInt8}, out, info, Xn, Un) = GC.@preserve X U ω Δ act out info Xn Un check(ctx, ccall((:gusto_iterate_host, LIB), Int32,
    (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{UInt8}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
    ctx.ptr, X, U, ω, Δ, act, out, info, Xn, Un))
# post-processing of the accepted trajectory (dynamics_constraint_satisfaction, verify_collision_free, interpolate_traj:
# dynamics/astrobee_se3.jl:495-562).  out is B x 8 (see GUSTO_CHECK_NOUT in the header); Xfull / Ufull are
# (x_dim, nstep*(N-1)+1, B) and (u_dim, nstep*(N-1), B).
check_trajectory!(ctx, out) = GC.@preserve out check(ctx, ccall((:gusto_check_trajectory, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), ctx.ptr, out))
interpolate_trajectory!(ctx, nstep::Integer, Xfull, Ufull) = GC.@preserve Xfull Ufull check(ctx, ccall((:gusto_interpolate_trajectory, LIB), Int32,
  (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}), ctx.ptr, Int32(nstep), Xfull, Ufull))
# shooting refinement (shooting.jl:4-66): duals is x_dim x B, out is 8 x B (GUSTO_SHOOT_NOUT), p0 / x_goal may be `nothing`
get_duals!(ctx, duals) = GC.@preserve duals check(ctx, ccall((:gusto_get_duals, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), ctx.ptr, duals))
function shoot!(ctx, p0, x_goal, out; nsub::Integer=4, max_iter::Integer=100, ftol::Float64=1e-3)
  pp = p0 === nothing ? Ptr{Float64}(C_NULL) : pointer(p0)
  pg = x_goal === nothing ? Ptr{Float64}(C_NULL) : pointer(x_goal)
  GC.@preserve p0 x_goal out check(ctx, ccall((:gusto_shoot, LIB), Int32,
    (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int32, Int32, Float64, Ptr{Float64}), ctx.ptr, pp, pg, Int32(nsub), Int32(max_iter), ftol, out))
end
get_shooting_trajectory!(ctx, X, U, P) = GC.@preserve X U P check(ctx, ccall((:gusto_get_shooting_trajectory, LIB), Int32,
    (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), ctx.ptr, X, U, P))
# device-resident outer loop and the status all-gather (include/gusto_b200.h)
const HIST_W = 12
scp_begin!(ctx, X, U, force::Bool) = GC.@preserve X U check(ctx, ccall((:gusto_scp_begin, LIB), Int32,
    (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int32), ctx.ptr, X, U, Int32(force)))
function scp_run!(ctx, max_iter::Integer)
  n = Ref{Int32}(0); u = Ref{Int32}(0)
  check(ctx, ccall((:gusto_scp_run, LIB), Int32, (Ptr{Cvoid}, Int32, Ref{Int32}, Ref{Int32}), ctx.ptr, Int32(max_iter), n, u))
  return Int(n[]), Int(u[])
end
scp_get!(ctx, iterations::Vector{Int32}, converged::Vector{UInt8}, successful::Vector{UInt8}, hist::Array{Float64,3}) =
  GC.@preserve iterations converged successful hist check(ctx, ccall((:gusto_scp_get, LIB), Int32,
    (Ptr{Cvoid}, Ptr{Int32}, Ptr{UInt8}, Ptr{UInt8}, Ptr{Float64}, Int32, Ptr{Int32}), ctx.ptr, iterations, converged, successful, hist,
    Int32(size(hist, 3)), Ptr{Int32}(C_NULL)))
function comm_unique_id()
  id = zeros(UInt8, 128)
  rc = ccall((:gusto_comm_unique_id, LIB), Int32, (Ptr{UInt8},), id)
  rc == 0 || error("gusto_comm_unique_id failed ($rc): " * unsafe_string(ccall((:gusto_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
  return id
end
comm_init!(ctx, rank::Integer, nranks::Integer, id::Vector{UInt8}) = GC.@preserve id check(ctx, ccall((:gusto_comm_init, LIB), Int32,
    (Ptr{Cvoid}, Int32, Int32, Ptr{UInt8}), ctx.ptr, Int32(rank), Int32(nranks), id))
function allgather_status!(ctx, done::Vector{UInt8})
  n = Ref{Int32}(0)
  GC.@preserve done check(ctx, ccall((:gusto_allgather_status, LIB), Int32, (Ptr{Cvoid}, Ptr{UInt8}, Ptr{UInt8}, Ref{Int32}),
    ctx.ptr, done, Ptr{UInt8}(C_NULL), n))
  return Int(n[])
end
accept!(ctx, acc::Vector{UInt8}, ω, Δ) = GC.@preserve acc ω Δ check(ctx, ccall((:gusto_accept, LIB), Int32,
    (Ptr{Cvoid}, Ptr{UInt8}, Ptr{Float64}, Ptr{Float64}), ctx.ptr, acc, ω, Δ))

# ------------------------------------------------------------------------------------- shooting plug-in
"""
    solve_shooting_b200!(SS, SP; device=0, nsub=4)

Same contract as `solve!(SS::ShootingSolution, SP::ShootingProblem)` (shooting.jl:4-49): one shooting attempt started
from `SP.p0` (= `SCPS.dual`); pushes `:Optimal` / `:Diverged`, `J_true`, `convergence_measure`, `iter_elapsed_times` and, on
success, replaces `SS.traj`.  DubinsCar and AstrobeeSE3Manifold only (the models for which the reference defines
`shooting_ode!`).
"""
function solve_shooting_b200!(SS, SP; device::Int=0, nsub::Int=4)
  model, robot, env = SP.PD.model, SP.PD.robot, SP.PD.env
  x_dim, u_dim, N = model.x_dim, model.u_dim, SP.N
  gtype = zeros(Int32, 16); gtype[1:x_dim] .= GOAL_POINT             # goal_type::NTuple{16,Int32}: padded like flatten_goals
  ctx = GustoContext(robot, model, env, N, 1, gtype, Main.SCPParam_GuSTO(model), (convergence_threshold = 0.0,); device=device)   # SCP parameters unused by K7
  xg = Float64.(SP.x_goal)
  set_problems!(ctx, Float64.(SP.PD.x_init), xg, xg, Float64[SP.tf])
  set_trajectory!(ctx, Matrix{Float64}(SS.traj.X), Matrix{Float64}(SS.traj.U))     # convergence_metric is taken against SS.traj
  out = zeros(8)
  time_start = time_ns()
  shoot!(ctx, Float64.(SP.p0), xg, out; nsub=nsub)
  iter_elapsed_time = (time_ns() - time_start)/10^9
  if out[1] == 0
    X = zeros(x_dim, N); U = zeros(u_dim, N); P = zeros(x_dim, N)
    get_shooting_trajectory!(ctx, X, U, P)
    push!(SS.prob_status, :Optimal); push!(SS.J_true, out[4]); push!(SS.convergence_measure, out[5])
    push!(SS.iter_elapsed_times, iter_elapsed_time)
    SS.traj.X = X; SS.traj.U = U; SS.traj.Tf = SP.tf; SS.traj.dt = SP.tf/(N-1)   # fresh arrays: what copy!(::Trajectory, ::Trajectory) does (types.jl:247-252)
  else
    push!(SS.prob_status, :Diverged); push!(SS.J_true, NaN); push!(SS.convergence_measure, NaN)
    push!(SS.iter_elapsed_times, iter_elapsed_time)
  end
  return
end

# ------------------------------------------------------------------------------------- single-instance plug-in
"""
    solve_gusto_b200!(SCPS, SCPP, solver="B200", max_iter=30, force=false; device=0, kwarg...)

Same contract as `solve_gusto_jump!` (scp_gusto.jl:49-176): mutates `SCPS.traj`, pushes to the SCPSolution histories
and to `SCPP.param.alg.{Δ_vec, ω_vec, ρ_vec, trust_region_satisfied_vec, convex_ineq_satisfied_vec}`, sets
`converged / successful / iterations / total_time`.  A solver failure pushes the elapsed time and returns with
`converged == false` (:107-111); ω > ω_max breaks out of the loop (:163-166).
"""
function solve_gusto_b200!(SCPS, SCPP, solver="B200", max_iter=30, force=false; device::Int=0, kwarg...)
  N = SCPP.N
  param, model, robot, env = SCPP.param, SCPP.PD.model, SCPP.PD.robot, SCPP.PD.env
  !isdefined(param, :alg) ? param.alg = Main.SCPParam_GuSTO(model) : nothing
  alg = param.alg
  Δ0, ω_max, ρ0, ρ1 = alg.Δ0, alg.ω_max, alg.ρ0, alg.ρ1
  β_succ, β_fail, γ_fail = alg.β_succ, alg.β_fail, alg.γ_fail
  x_dim, u_dim = model.x_dim, model.u_dim
  gtype, glo, ghi = flatten_goals(SCPP.PD.goal_set, x_dim, SCPP.tf_guess)

  ctx = GustoContext(robot, model, env, N, 1, gtype, alg, param; device=device)
  set_problems!(ctx, Float64.(SCPP.PD.x_init), glo, ghi, Float64[SCPS.traj.Tf])
  X = Matrix{Float64}(SCPS.traj.X); U = Matrix{Float64}(SCPS.traj.U)      # x_dim x N column-major == [N][x_dim] knot-major
  set_trajectory!(ctx, X, U); set_candidate!(ctx, X, U)
  set_penalties!(ctx, Float64[alg.ω_vec[end]], Float64[alg.Δ_vec[end]])
  out = zeros(EVAL_NOUT); info = zeros(SOLVE_NINFO)

  iter_cap = SCPS.iterations + max_iter
  linearize!(ctx); evaluate!(ctx, out)                                     # :72-75
  push!(SCPS.J_true, out[5]); push!(SCPS.J_full, SCPS.J_true[end]); push!(alg.ρ_vec, out[4])
  param.obstacle_toggle_distance = alg.Δ_vec[end]/8 + model.clearance

  while SCPS.iterations < iter_cap
    time_start = time_ns()
    iterate!(ctx, out, info)                                               # K1+K2, K3, K4
    push!(SCPS.solver_status, solver_status_symbol(info[1]))
    if !(info[1] == 0 || info[1] == 3)                                     # :107-111 (OPTIMAL / ALMOST_OPTIMAL continue)
      push!(SCPS.iter_elapsed_times, (time_ns() - time_start)/10^9)
      break                                                                # SCPS.traj = last accepted iterate, copied below
    end
    conv, tr_ok, ineq_ok, ρ, J_new, J_full = out[1], out[2] > 0.5, out[3] > 0.5, out[4], out[5], info[5]
    push!(SCPS.convergence_measure, conv); push!(SCPS.J_full, J_full)
    push!(alg.trust_region_satisfied_vec, tr_ok); push!(alg.convex_ineq_satisfied_vec, ineq_ok)
    Δ, ω = alg.Δ_vec[end], alg.ω_vec[end]
    if tr_ok
      push!(alg.ρ_vec, ρ)
      if ρ > ρ1
        push!(SCPS.scp_status, :InaccurateModel); push!(SCPS.accept_solution, false)
        push!(alg.Δ_vec, β_fail*Δ); push!(alg.ω_vec, ω)
      else
        push!(SCPS.accept_solution, true)
        ρ < ρ0 ? push!(alg.Δ_vec, min(β_succ*Δ, Δ0)) : push!(alg.Δ_vec, Δ)
        if !ineq_ok
          push!(SCPS.scp_status, :ViolatesConstraints); push!(alg.ω_vec, γ_fail*ω)
        else
          push!(SCPS.scp_status, :OK); push!(alg.ω_vec, ω)
        end
      end
    else
      push!(SCPS.scp_status, :TrustRegionViolated); push!(SCPS.accept_solution, false)
      push!(alg.Δ_vec, Δ); push!(alg.ω_vec, γ_fail*ω)
    end
    accept!(ctx, UInt8[SCPS.accept_solution[end]], Float64[alg.ω_vec[end]], Float64[alg.Δ_vec[end]])
    push!(SCPS.J_true, SCPS.accept_solution[end] ? J_new : SCPS.J_true[end])
    param.obstacle_toggle_distance = alg.Δ_vec[end]/8 + model.clearance

    iter_elapsed_time = (time_ns() - time_start)/10^9
    push!(SCPS.iter_elapsed_times, iter_elapsed_time)
    SCPS.total_time += iter_elapsed_time
    SCPS.iterations += 1

    if alg.ω_vec[end] > ω_max
      @warn "GuSTO SCP omegamax exceeded"
      break
    end
    !SCPS.accept_solution[end] ? continue : nothing
    conv_iter_spread = 2
    if SCPS.iterations > conv_iter_spread && sum(SCPS.convergence_measure[end-conv_iter_spread+1:end]) <= param.convergence_threshold
      SCPS.converged = true
      alg.convex_ineq_satisfied_vec[end] && (SCPS.successful = true)
      force ? continue : break
    end
  end
  # every exit path (converged, iteration cap, omega_max, solver failure) leaves the last ACCEPTED iterate in SCPS.traj, as the
  # reference does by copying on every accepted iteration (:147); X, U are fresh arrays, which is what copy!(::Trajectory,
  # ::Trajectory) assigns (types.jl:247-252: a.X = deepcopy(b.X))
  get_trajectory!(ctx, X, U)
  SCPS.traj.X = X; SCPS.traj.U = U
  SCPS.dual = zeros(x_dim); get_duals!(ctx, SCPS.dual)     # -JuMP.dual of the init constraints (:116, get_dual_jump): p0 of the shooting refinement
  return
end

# ------------------------------------------------------------------------------------- TrajOpt plug-in (SURVEY 8(f)-1)
const TRAJOPT_NOUT = 8
trajopt_enable!(ctx) = check(ctx, ccall((:gusto_trajopt_enable, LIB), Int32, (Ptr{Cvoid},), ctx.ptr))
trajopt_iterate!(ctx, μ, s, act::Vector{UInt8}, out, info) = GC.@preserve μ s act out info check(ctx, ccall((:gusto_trajopt_iterate, LIB), Int32,
    (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{UInt8}, Ptr{Float64}, Ptr{Float64}), ctx.ptr, μ, s, act, out, info))
trajopt_mark!(ctx, slot::Integer) = check(ctx, ccall((:gusto_trajopt_mark, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{UInt8}), ctx.ptr, Int32(slot), C_NULL))
trajopt_compare!(ctx, slot::Integer, out) = GC.@preserve out check(ctx, ccall((:gusto_trajopt_compare, LIB), Int32,
    (Ptr{Cvoid}, Int32, Ptr{Float64}), ctx.ptr, Int32(slot), out))

"""
    solve_trajopt_b200!(SCPS, SCPP, solver="B200", max_iter=125, force=false; device=0, kwarg...)

Same slot and contract as `solve_trajopt_jump!` (scp_trajopt.jl:33-157): the three nested loops (penalty / convex / trust) stay here,
every convex subproblem (:159-279) and every evaluation scalar runs on the device.  The reference routine cannot run as written; this
restates it with the repairs listed in oracle/gusto_oracle/trajopt.py (bounded l1 dynamics penalty, real copies for
old_penalty_traj / old_convex_traj after the trust loop, `xtol_vec[end]` in :142, evaluate_ctol over the model's constraint classes).
Un-executed here (no Julia in the build image); gusto.jl_b200/host.py::solve_trajopt_batch is the tested twin.
"""
function solve_trajopt_b200!(SCPS, SCPP, solver="B200", max_iter=125, force=false; device::Int=0, kwarg...)
  N = SCPP.N
  param, model, robot, env = SCPP.param, SCPP.PD.model, SCPP.PD.robot, SCPP.PD.env
  param.alg = Main.SCPParam_TrajOpt(model)                                   # :44
  alg = param.alg
  x_dim, u_dim = model.x_dim, model.u_dim
  gtype, glo, ghi = flatten_goals(SCPP.PD.goal_set, x_dim, SCPP.tf_guess)
  gusto_alg = Main.SCPParam_GuSTO(model)                                     # only fills the configuration slots TrajOpt does not read
  ctx = GustoContext(robot, model, env, N, 1, gtype, gusto_alg, param; device=device)
  trajopt_enable!(ctx)
  set_problems!(ctx, Float64.(SCPP.PD.x_init), glo, ghi, Float64[SCPS.traj.Tf])
  X = Matrix{Float64}(SCPS.traj.X); U = Matrix{Float64}(SCPS.traj.U)
  set_trajectory!(ctx, X, U); set_candidate!(ctx, X, U)
  ev = zeros(EVAL_NOUT); out = zeros(TRAJOPT_NOUT); info = zeros(SOLVE_NINFO); cmp = zeros(5); live = UInt8[1]
  linearize!(ctx); evaluate!(ctx, ev)
  push!(SCPS.J_true, ev[5])                                                  # :64
  param.obstacle_toggle_distance = model.clearance + 1.                      # :65 (the kernel applies the same constant)
  constraints_satisfied = false; xtol_satisfied = false; failed = false
  for penalty_iteration in 1:alg.max_penalty_iteration
    (constraints_satisfied || failed) && break
    trajopt_mark!(ctx, 0)                                                    # :73
    for convex_iteration in 1:alg.max_convex_iteration
      trajopt_mark!(ctx, 1)                                                  # :76
      (constraints_satisfied || failed) && break
      if xtol_satisfied
        xtol_satisfied = false
        break
      end
      for trust_iteration in 1:alg.max_trust_iteration
        time_start = time_ns()
        trajopt_iterate!(ctx, Float64[alg.mu_vec[end]], Float64[alg.s_vec[end]], live, out, info)
        push!(SCPS.solver_status, solver_status_symbol(info[1]))
        if !(info[1] == 0 || info[1] == 3)                                   # :107-111 warns and goes on; there is no iterate to go on with
          push!(SCPS.iter_elapsed_times, (time_ns() - time_start)/10^9)
          failed = true
          break
        end
        push!(alg.xtol_vec, out[1]); push!(SCPS.convergence_measure, out[1]); push!(SCPS.J_full, info[5])
        push!(alg.ρ_vec, out[2])
        push!(alg.s_vec, (out[2] > alg.c ? alg.τ_plus : alg.τ_minus)*alg.s_vec[end])      # :122-126
        accept!(ctx, live, Float64[alg.mu_vec[end]], Float64[alg.s_vec[end]])            # copy!(SCPS.traj, new_traj), :128
        iter_elapsed_time = (time_ns() - time_start)/10^9
        push!(SCPS.J_true, out[3]); push!(SCPS.iter_elapsed_times, iter_elapsed_time)
        SCPS.total_time += iter_elapsed_time
        SCPS.iterations += 1
        if alg.s_vec[end] < alg.xtol
          xtol_satisfied = true
          break
        end
      end
      failed && break
      trajopt_compare!(ctx, 1, cmp)                                          # :140-141
      push!(alg.ftol_vec, abs(cmp[4] - cmp[5])/abs(cmp[4])); push!(alg.xtol_vec, cmp[3])
      if alg.ftol_vec[end] < alg.ftol || alg.xtol_vec[end] < alg.xtol
        constraints_satisfied = true
        break
      end
    end
    failed && break
    trajopt_compare!(ctx, 0, cmp)                                            # :148
    push!(alg.ctol_vec, cmp[1]/cmp[2])
    if alg.ctol_vec[end] < alg.ctol
      constraints_satisfied = true
      SCPS.converged = true
      break
    else
      push!(alg.mu_vec, alg.mu_vec[end]*alg.k)
    end
  end
  get_trajectory!(ctx, X, U)
  SCPS.traj.X = X; SCPS.traj.U = U
  SCPS.dual = zeros(x_dim); get_duals!(ctx, SCPS.dual)
  return
end

# ------------------------------------------------------------------------------------------------ batched driver
"""
    solve_SCP_batch!(TOPs, init_method; max_iter=30, force=false, device=0) -> (X, U, converged, successful, iterations)

B TrajectoryOptimizationProblems that share robot / model / environment / N / goal coordinates, solved together.
Returns X (x_dim, N, B), U (u_dim, N, B) and per-instance flags; the per-instance logic is the one of
`solve_gusto_b200!` vectorised over B (see gusto.jl_b200/host.py::gusto_update, which is the tested twin).
"""
function solve_SCP_batch!(TOPs::Vector, init_method; max_iter::Int=30, force::Bool=false, device::Int=0)
  B = length(TOPs); T1 = TOPs[1]
  model, robot, env, N = T1.PD.model, T1.PD.robot, T1.PD.env, T1.N
  x_dim, u_dim = model.x_dim, model.u_dim
  alg = Main.SCPParam_GuSTO(model); param = Main.SCPParam(model, T1.fixed_final_time)
  gtype, _, _ = flatten_goals(T1.PD.goal_set, x_dim, T1.tf_guess)
  x_init = zeros(x_dim, B); glo = zeros(x_dim, B); ghi = zeros(x_dim, B); tf = zeros(B)
  X = zeros(x_dim, N, B); U = zeros(u_dim, N, B)
  for (b, TOP) in enumerate(TOPs)
    _, lo, hi = flatten_goals(TOP.PD.goal_set, x_dim, TOP.tf_guess)
    x_init[:, b] = TOP.PD.x_init; glo[:, b] = lo; ghi[:, b] = hi; tf[b] = TOP.tf_guess
    traj = init_method(TOP); X[:, :, b] = traj.X; U[:, :, b] = traj.U
  end
  ctx = GustoContext(robot, model, env, N, B, gtype, alg, param; device=device)
  set_problems!(ctx, x_init, glo, ghi, tf)
  set_trajectory!(ctx, X, U)
  Δ = fill(alg.Δ0, B); ω = fill(alg.ω0, B); set_penalties!(ctx, ω, Δ)
  out = zeros(EVAL_NOUT, B); info = zeros(SOLVE_NINFO, B)
  active = trues(B); converged = falses(B); successful = falses(B); iterations = zeros(Int, B); conv_prev = zeros(B)
  for it in 1:max_iter
    set_active!(ctx, UInt8.(active))
    iterate!(ctx, out, info)
    acc = zeros(UInt8, B)
    for b in 1:B
      active[b] || continue
      if !(info[1, b] == 0 || info[1, b] == 3); active[b] = false; continue; end            # :107-111
      conv, tr_ok, ineq_ok, ρ = out[1, b], out[2, b] > 0.5, out[3, b] > 0.5, out[4, b]
      accepted = false
      if tr_ok
        if ρ > alg.ρ1
          Δ[b] *= alg.β_fail
        else
          accepted = true
          ρ < alg.ρ0 && (Δ[b] = min(alg.β_succ*Δ[b], alg.Δ0))
          !ineq_ok && (ω[b] *= alg.γ_fail)
        end
      else
        ω[b] *= alg.γ_fail
      end
      acc[b] = accepted; iterations[b] += 1
      if ω[b] > alg.ω_max
        active[b] = false
      elseif accepted && iterations[b] > 2 && conv + conv_prev[b] <= param.convergence_threshold
        converged[b] = true; successful[b] = ineq_ok
        force || (active[b] = false)
      end
      conv_prev[b] = conv
    end
    accept!(ctx, acc, ω, Δ)
    # one status all-gather per outer iteration (in-library NCCL when a communicator was set up with comm_init!; the local
    # count otherwise): every rank leaves the loop in the same iteration
    allgather_status!(ctx, UInt8.(.!active)) == 0 && break
  end
  get_trajectory!(ctx, X, U)
  return X, U, converged, successful, iterations
end

"""
    solve_SCP_batch_device!(TOPs, init_method; max_iter=30, force=false, device=0, rank=0, nranks=1, comm_id=nothing)

The same batched solve with the outer loop resident on the device (`gusto_scp_begin` / `gusto_scp_run`, include/gusto_b200.h):
accept / reject, the Δ / ω schedule and the convergence test of scp_gusto.jl:119-174 run in the library's update kernel, one
CUDA graph per outer iteration, and the host reads one counter per iteration.  Multi-GPU: one Julia process per GPU, each with
its shard of `TOPs`; rank 0 calls `comm_unique_id()` and ships the 128 bytes to the others (MPI.Bcast!, a file, ...), every rank
passes them as `comm_id` together with `rank` / `nranks`; the library then all-gathers the status bytes over NCCL once per
iteration and all ranks stop together.  Returns (X, U, converged, successful, iterations, hist) with
hist (12, B, batch_iterations + 1): the per-iteration records described at `gusto_scp_get`.
"""
function solve_SCP_batch_device!(TOPs::Vector, init_method; max_iter::Int=30, force::Bool=false, device::Int=0,
                                 rank::Int=0, nranks::Int=1, comm_id=nothing)
  B = length(TOPs); T1 = TOPs[1]
  model, robot, env, N = T1.PD.model, T1.PD.robot, T1.PD.env, T1.N
  x_dim, u_dim = model.x_dim, model.u_dim
  alg = Main.SCPParam_GuSTO(model); param = Main.SCPParam(model, T1.fixed_final_time)
  gtype, _, _ = flatten_goals(T1.PD.goal_set, x_dim, T1.tf_guess)
  x_init = zeros(x_dim, B); glo = zeros(x_dim, B); ghi = zeros(x_dim, B); tf = zeros(B)
  X = zeros(x_dim, N, B); U = zeros(u_dim, N, B)
  for (b, TOP) in enumerate(TOPs)
    _, lo, hi = flatten_goals(TOP.PD.goal_set, x_dim, TOP.tf_guess)
    x_init[:, b] = TOP.PD.x_init; glo[:, b] = lo; ghi[:, b] = hi; tf[b] = TOP.tf_guess
    traj = init_method(TOP); X[:, :, b] = traj.X; U[:, :, b] = traj.U
  end
  ctx = GustoContext(robot, model, env, N, B, gtype, alg, param; device=device)
  set_problems!(ctx, x_init, glo, ghi, tf)
  nranks > 1 && comm_init!(ctx, rank, nranks, comm_id)
  scp_begin!(ctx, X, U, force)
  n_it, _ = scp_run!(ctx, max_iter)
  iterations = zeros(Int32, B); converged = zeros(UInt8, B); successful = zeros(UInt8, B); hist = zeros(HIST_W, B, n_it + 1)
  scp_get!(ctx, iterations, converged, successful, hist)
  get_trajectory!(ctx, X, U)
  return X, U, converged .!= 0, successful .!= 0, Int.(iterations), hist
end

# ------------------------------------------------------------------------------------- BulletCollision stand-in
# The unchanged reference structs call BulletCollision at construction time (robot/astrobee3D.jl:30, types.jl:12-24)
# although the B200 path never queries it.  On a Julia without Cxx.jl/BulletCollision.jl, load this stub first:
#     include("GuSTOB200.jl"); const BulletCollision = GuSTOB200.BulletCollisionStub
module BulletCollisionStub
  struct BulletCollisionObjectPtr; kind::Symbol; data; end
  struct BulletStaticEnvironment
    robot; world
    convex_env_components::Vector{Any}
    convex_robot_components::Vector{Any}
  end
  sphere(c, r) = BulletCollisionObjectPtr(:sphere, (c, r))
  convex_hull(pts) = BulletCollisionObjectPtr(:hull, pts)
  convex_hull_cylinder(a, b, r) = BulletCollisionObjectPtr(:cylinder, (a, b, r))
  compound_collision_object(objs) = BulletCollisionObjectPtr(:compound, objs)
  collision_world(lo, hi) = (lo, hi)
  geometry_type_to_BT(zone) = BulletCollisionObjectPtr(:zone, zone)
  BulletStaticEnvironment(robot, world) = BulletStaticEnvironment(robot, world, Any[],
      robot.kind == :compound ? Any[robot.data...] : Any[robot])
  add_collision_object!(env::BulletStaticEnvironment, obj) = push!(env.convex_env_components, obj)
  distance(args...) = error("BulletCollisionStub.distance: signed distances are computed on the GPU (csrc/sdf.cuh)")
end

end # module
