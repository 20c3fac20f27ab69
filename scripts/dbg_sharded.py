"""Debug aid (torchrun, >= 2 GPUs): sharded pass with the compact vs the dense linearisation, and solve_sharded with
and without graph replay, against the single-GPU solve."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "parallel-in-time-ode-filters_b200")]
import numpy as np, torch, torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import pof.ivp
from pof import _native as nat
from pof.convenience import get_initial_trajectory, set_up_solver
from pof.sharded import ShardedPass, shard_bounds, solve_sharded
from pof.solver import solve
from pof.step import linearize_at_previous_states

ivp = pof.ivp.rigid_body()
N, q, d = 20000, 3, 3
D = d * (q + 1)
ts = np.linspace(ivp.t0, ivp.tmax, N)
setup = set_up_solver(f=ivp.f, y0=ivp.y0, ts=ts, order=q)
lin = setup["om"].f._pof_lin
st = get_initial_trajectory(setup, method="constant")
dom = linearize_at_previous_states(setup["om"], st)
k_lo, k_hi = shard_bounds(N - 1, rank, world)
r0 = 0 if rank == 0 else k_lo + 1
sp = ShardedPass(N, d, q, setup["_qL"], rank=rank, world=world, device=dev)
outs = []
for mode in ("dense", "compact"):
    means = st.mean[r0:k_hi + 1].contiguous().clone()
    chols = torch.zeros((sp.rows, D, D), dtype=torch.float64, device=dev)
    if mode == "dense":
        res = sp.run(setup["x0"].mean, setup["x0"].chol, dom.H[k_lo:k_hi].contiguous(), dom.b[k_lo:k_hi].contiguous(),
                     means, chols, calibrate=True)
    else:
        sp.backend.set_compact(lin["scale0"], lin["scale1"])
        Jc = torch.empty((sp.n_loc, d * d + d), dtype=torch.float64, device=dev)
        ivp_id, params = lin["builtin"]
        ph, pp = nat.host_doubles(list(params) + [0.0])
        m0 = st.mean[r0:k_hi + 1].contiguous()
        t1row = 1 if rank == 0 else 0
        nat.check(nat.LIB.pof_linearize_ivp_compact_f64(nat.stream_ptr(), ivp_id, pp, len(params), sp.n_loc, d, q,
                                                        lin["scale0"], nat.ptr(m0[t1row:]), nat.ptr(Jc)), "lin")
        res = sp.run(setup["x0"].mean, setup["x0"].chol, Jc, None, means, chols, calibrate=True)
    torch.cuda.synchronize()
    outs.append((means.clone(), {k: float(v) for k, v in res.items()}))
dm = float((outs[0][0] - outs[1][0]).abs().max() / outs[0][0].abs().max())
print(f"[rank {rank}] one pass dense vs compact: rel diff {dm:.3e}; dense {outs[0][1]}; compact {outs[1][1]}", flush=True)
ref, rinfo = solve(f=ivp.f, y0=ivp.y0, ts=ts, order=q, init="constant", maxiters=200)
for g in (False, True):
    ys, info, rows = solve_sharded(f=ivp.f, y0=ivp.y0, ts=ts, order=q, init="constant", maxiters=200, graph=g)
    em = float((ys.mean - ref.mean[rows]).abs().max())
    print(f"[rank {rank}] solve_sharded graph={g}: iterations {info['iterations']} (single GPU {rinfo['iterations']}) "
          f"max|dy| {em:.3e} obj {info['obj']:.6e} (single {rinfo['obj']:.6e})", flush=True)
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0)
