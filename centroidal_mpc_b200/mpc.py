"""Receding-horizon driver: B independent MPC loops that slide a horizon of H
knots along the gait of ``conf`` and re-solve the SCP problem at every tick.

The reference names receding-horizon operation but ships no driver; its pieces are all here: every tick is one
``solve_scp`` (/root/reference/src/scp_solver.py:118-179) of a ``Centroidal_model`` whose warm start
(``centroidal_traj`` / ``_init_trajectories``, src/centroidal_model.py:87,150-176) is the previous tick's solution
shifted by one knot, whose contact data are the window [k0, k0 + H) of the contact plan (src/centroidal_model.py:
132-148) and whose initial state is the state the robot reached.  Because the warm start is a solution, the
friction / CoP rows it sits on are a good guess of the new active set: with ``warm_start`` the device starts the
QP with the certified polish on that set and runs no ADMM iteration unless the certificate fails
(cmpc_qp_settings.warm_start, include/cmpc.h).

Everything between two ticks stays on the device: the shift is a copy between device tensors (torch as the batch
container), the contact window is a pointer offset into the gait's contact arrays."""
import ctypes as C

import numpy as np

from . import _lib as L
from . import synthetic
from .device import BatchSolver, _ptr, _torch_cuda


def shift_window(X, U, x_tail, u_tail, x_init=None):
    """One tick of the receding horizon on (torch or numpy) arrays: X [B,H+1,9], U [B,H,nu] are the solution of
    the previous tick, x_tail [B,9] the reference state that enters the horizon, u_tail [B,nu] the warm-start
    control of the knot that enters.  Returns (x_init, x_final, X_ref, U_init) of the next problem; x_init defaults
    to the planned next state X[:,1]."""
    cat = np.concatenate if isinstance(X, np.ndarray) else __import__("torch").cat
    X_ref = cat([X[:, 1:], x_tail[:, None, :]], 1)
    U_init = cat([U[:, 1:], u_tail[:, None, :]], 1)
    if x_init is not None:
        X_ref = X_ref.copy() if isinstance(X_ref, np.ndarray) else X_ref.clone()
        X_ref[:, 0] = x_init
    return X_ref[:, 0], X_ref[:, -1], X_ref, U_init


class RecedingHorizonMPC:
    def __init__(self, name, B, horizon, mode="B", warm=None, first=0):
        """``warm``: start every tick after the first with the certified polish on the active set of the shifted
        previous solution (cmpc_qp_settings.warm_start).  Default: on for the wrench contact model (talos: 8.7 ms
        per tick instead of 10.2 ms, 4096 loops, H = 100, B200), off for point contacts (solo12 trot: 14.4 ms
        against 11.2 ms cold -- there the previous rows are a worse guess than eight ADMM iterations)."""
        torch = _torch_cuda()
        self.conf_full = synthetic.load_conf(name)
        self.conf = synthetic.load_conf(name, N=horizon)
        self.H, self.B = int(horizon), int(B)
        self.warm = (getattr(self.conf, "robot_name", "") == "TALOS") if warm is None else bool(warm)
        self.Ntot = int(self.conf_full.N)
        if self.H >= self.Ntot:
            raise ValueError("the horizon must be shorter than the gait (%d knots)" % self.Ntot)
        full = synthetic.make_batch(self.conf_full, B, mode=mode, first=first)      # the whole gait: references, plan
        if not full.shared_plan:
            raise ValueError("the driver slides ONE contact plan under all instances")
        win = synthetic.make_batch(self.conf, B, mode=mode, first=first)            # the first window
        self.solver = BatchSolver(win)
        dev = self.solver.device
        self.X_gait = torch.from_numpy(full.X_ref).to(dev)                           # [B, Ntot+1, 9]
        self.U_gait = torch.from_numpy(full.U_init).to(dev)                          # [B, Ntot, nu] default warm start
        self.cpos = torch.from_numpy(full.contact_pos).to(dev)                       # [1, Ntot, nc, 3]
        self.cact = torch.from_numpy(full.contact_active).to(dev)
        self.cR = None if full.contact_R is None else torch.from_numpy(full.contact_R).to(dev)
        self.k0 = 0
        self.ticks = 0
        self.failed = torch.zeros(B, dtype=torch.bool, device=dev)
        self._bind(self.solver.d["x_init"], self.solver.d["x_final"], self.solver.d["X_ref"], self.solver.d["U_init"])

    def _bind(self, x_init, x_final, X_ref, U_init):
        """Problem pointers of the next solve: the contact arrays are the gait's, offset to the window."""
        s = self.solver
        self._keep = (x_init.contiguous(), x_final.contiguous(), X_ref.contiguous(), U_init.contiguous())
        nc = self.cact.shape[2]
        off = lambda t, per_knot: None if t is None else C.c_void_p(t.data_ptr() + self.k0 * per_knot * t.element_size())
        L.check(s.lib.cmpc_set_problem(s.handle, C.byref(s.model), _ptr(self._keep[0]), _ptr(self._keep[1]), _ptr(self._keep[2]),
                                       _ptr(self._keep[3]), off(self.cpos, nc * 3), off(self.cR, nc * 9), off(self.cact, nc)), s.lib)

    def step(self, disturbance=None, qp_overrides=None):
        """Solve the current window, then slide it by one knot.  Returns (u0 [B,nu], status [B]) as device tensors:
        the control to apply and the per-instance solver status.  ``disturbance`` [B,9] is added to the planned next
        state (the state the robot actually reached)."""
        torch = _torch_cuda()
        s = self.solver
        ov = dict(qp_overrides or {})
        if self.warm and self.ticks > 0:
            ov["warm_start"] = 1
        s.solve(self.conf.scp_params, ov or None)
        X, U, status, nacc = s.X, s.U, s.ints[1], s.ints[2]
        ok = (status == 0) & (nacc > 0)
        self.failed |= ~ok
        u0 = U[:, 0].clone()
        if self.k0 + self.H + 1 > self.Ntot:
            raise StopIteration("the window reached the end of the gait")
        self.k0 += 1
        k1 = self.k0 + self.H
        x_next = X[:, 1] if disturbance is None else X[:, 1] + disturbance
        # an instance whose solve failed keeps following the gait's own reference / default warm start
        Xs = torch.where(ok[:, None, None], X, self._keep[2])
        Us = torch.where(ok[:, None, None], U, self._keep[3])
        x_init, x_final, X_ref, U_init = shift_window(Xs, Us, self.X_gait[:, k1], self.U_gait[:, k1 - 1], x_init=x_next)
        self._bind(x_init, x_final, X_ref, U_init)
        self.ticks += 1
        return u0, status

    def stats(self):
        return self.solver.stats()

    def close(self):
        self.solver.close()
