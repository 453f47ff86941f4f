"""Batch container: packs independent MPC instances into the instance-major arrays of the
C ABI (include/cmpc.h).  Pure layout code (numpy); no arithmetic of the hot path lives here."""
import numpy as np


class ProblemBatch:
    """B problems that share dims, robot constants and cost weights (one config) and differ in
    x_init / x_final / reference trajectories / warm-start controls, optionally in the contact
    plan.  Build it from ``Centroidal_model`` objects or from ``problem_arrays()`` dicts."""

    def __init__(self, probs, shared_plan=None):
        probs = [p.problem_arrays() if hasattr(p, "problem_arrays") else p for p in probs]
        if not probs:
            raise ValueError("empty batch")
        p0 = probs[0]
        self.B = len(probs)
        self.N = int(p0["N"])
        self.nc = int(p0["contact_active"].shape[1])
        # TALOS: flat feet, six controls (cop_x, cop_y, fx, fy, fz, tau_z) per foot -- the wrench contact model
        self.wrench = p0["robot"] == "TALOS"
        if self.wrench and self.nc > 2:
            raise ValueError("the wrench contact model takes at most two feet")
        if self.wrench and p0.get("stochastic") is not None:
            raise NotImplementedError("STOCHASTIC_OCP is not available for the wrench contact model")
        self.nu = (6 if self.wrench else 3) * self.nc
        self.proto = p0
        for p in probs[1:]:
            if int(p["N"]) != self.N or p["contact_active"].shape[1] != self.nc:
                raise ValueError("all instances of a batch must share N and the number of contacts")
            for key in ("m", "g", "dt", "mu", "robot"):
                if p[key] != p0[key]:
                    raise ValueError("all instances of a batch must share %s" % key)
            # the batch is solved with ONE model struct (weights) and one mode (nominal / stochastic)
            for key in ("state_cost_weights", "control_cost_weights"):
                if not np.array_equal(np.asarray(p[key]), np.asarray(p0[key])):
                    raise ValueError("all instances of a batch must share %s" % key)
            s0, s1 = p0.get("stochastic"), p.get("stochastic")
            if (s0 is None) != (s1 is None):
                raise ValueError("a batch cannot mix nominal and STOCHASTIC_OCP models")
            if s0 is not None and any(not np.array_equal(np.asarray(s0[k]), np.asarray(s1[k])) for k in s0):
                raise ValueError("all instances of a stochastic batch must share beta_u, Q, R, cov_w, cov_eta")
        if shared_plan is None:
            shared_plan = all(p["contact_active"] is p0["contact_active"] or
                              (np.array_equal(p["contact_active"], p0["contact_active"])
                               and np.array_equal(p["contact_pos"], p0["contact_pos"])
                               and np.array_equal(p["contact_R"], p0["contact_R"])) for p in probs[1:])
        self.shared_plan = bool(shared_plan)
        f64 = np.float64
        self.x_init = np.ascontiguousarray(np.stack([p["x_init"] for p in probs]), dtype=f64)
        self.x_final = np.ascontiguousarray(np.stack([p["x_final"] for p in probs]), dtype=f64)
        self.X_ref = np.ascontiguousarray(np.stack([p["X_ref"].T for p in probs]), dtype=f64)
        self.U_init = np.ascontiguousarray(np.stack([p["U_init"].T for p in probs]), dtype=f64)
        plan = probs[:1] if self.shared_plan else probs
        self.contact_pos = np.ascontiguousarray(np.stack([p["contact_pos"] for p in plan]), dtype=f64)
        self.contact_active = np.ascontiguousarray(np.stack([p["contact_active"] for p in plan]), dtype=np.int32)
        R = np.stack([p["contact_R"] for p in plan]).astype(f64)
        eye = np.eye(3)[None, None, None] * self.contact_active[..., None, None]
        self.identity_R = bool(np.array_equal(R, eye)) and not self.wrench   # the wrench model always reads the frames
        self.contact_R = None if self.identity_R else np.ascontiguousarray(R)

    @classmethod
    def from_arrays(cls, proto, x_init, x_final, X_ref, U_init):
        """Shared-plan batch from stacked arrays: X_ref [B,N+1,9], U_init [B,N,nu]."""
        self = cls([proto], shared_plan=True)
        f64 = np.float64
        self.B = int(X_ref.shape[0])
        self.x_init = np.ascontiguousarray(x_init, dtype=f64)
        self.x_final = np.ascontiguousarray(x_final, dtype=f64)
        self.X_ref = np.ascontiguousarray(X_ref, dtype=f64)
        self.U_init = np.ascontiguousarray(U_init, dtype=f64)
        return self

    def input_bytes(self):
        n = self.x_init.nbytes + self.x_final.nbytes + self.X_ref.nbytes + self.U_init.nbytes
        n += self.contact_pos.nbytes + self.contact_active.nbytes
        if self.contact_R is not None:
            n += self.contact_R.nbytes
        return n

    def output_bytes(self):
        return self.B * ((self.N + 1) * 9 + self.N * self.nu) * 8 + self.B * 3 * 4
