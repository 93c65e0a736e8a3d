"""Test double for stochvolmodels_b200.multi_gpu.CudaMcEngine: same interface, arithmetic delegated to the numpy oracle on CPU
tensors.  It lets the N>1 host logic (path sharding, the two all-reduces per maturity, finalisation, result assembly) run under
gloo without a GPU -- the analogue of the reference's fake pricers in tests/test_model_calibration_contracts.py:32-77."""
import numpy as np
import torch

from oracle import mc
from stochvolmodels_b200 import _capi as C


class OracleEngine:
    def __init__(self, model, params_c, n_local, path_offset, flags, max_strikes, device=None, factors=None):
        self.model, self.p, self.n_local, self.path_offset, self.flags = model, params_c, n_local, path_offset, flags
        self.factors = factors
        self.x = self.v = self.q = None
        self.moments = torch.zeros(2, dtype=torch.float64)
        self.device = torch.device("cpu")

    def to_device(self, array, dtype):
        return torch.as_tensor(np.ascontiguousarray(array), dtype=dtype)

    def simulate_slice(self, m, init, nsteps, dt, eta, is_spot, forward, seed):
        n = self.n_local
        ids = self.path_offset + np.arange(n, dtype=np.uint64)
        Z0, Z1 = mc.device_normals(seed, ids, 0 if self.model == "rough" else m, nsteps, "f64" if self.flags & C.GAUSS_F64 else "f32")
        if self.model == "rough":          # every maturity from t = 0 on its own grid, slice 0 of the path's stream (oracle/rough.py)
            from oracle import rough
            w, nodes = (np.asarray(a, dtype=float) for a in self.factors)
            volvol = np.sqrt(self.p.beta ** 2 + self.p.volvol ** 2)
            v0 = np.repeat(np.full((w.size,), self.p.sigma0 / w.sum())[:, None], n, axis=1)
            grid = dt * np.arange(nsteps + 1)
            grid[1] = dt                      # h = grid[1] - grid[0] exactly
            ls, _, qv = rough.log_spot_full_combined(np.repeat(nodes[:, None], n, axis=1), np.repeat(w[:, None], n, axis=1), v0, self.p.theta,
                                                     self.p.kappa1, self.p.kappa2, 0.0, v0.copy(), self.p.beta / volvol, volvol, grid, Z0, Z1)
            self.x, self.q = ls[0], qv[0]
        elif self.model == "logsv":
            if init:
                self.x, self.v, self.q = np.zeros(n), self.p.sigma0 * np.ones(n), np.zeros(n)
            self.x, self.v, self.q = mc.logsv_step_fixed(self.x, self.v, self.q, Z0, Z1, dt, self.p.theta, self.p.kappa1, self.p.kappa2,
                                                         self.p.beta, self.p.volvol, eta, is_spot)
        else:
            if init:
                self.x, self.v, self.q = np.zeros(n), self.p.v0 * np.ones(n), np.zeros(n)
            self.x, self.v, self.q = mc.heston_step_fixed(self.x, self.v, self.q, Z0, Z1, dt, self.p.theta, self.p.kappa, self.p.rho, self.p.volvol)
        spots = forward * np.exp(self.x)
        ok = ~np.isnan(spots)
        self.moments[0], self.moments[1] = float(spots[ok].sum()), float(ok.sum())
        return self.moments

    def payoff_sums(self, ttm, forward, strikes_dev, types_dev, J, variable_type, kinds=0):
        corr = float(self.moments[0] / self.moments[1]) - forward       # GLOBAL moments (all-reduced in place by the caller)
        spots = forward * np.exp(self.x) - corr
        under = spots if variable_type == C.LOG_RETURN else self.q / ttm
        out = torch.zeros(3 * J, dtype=torch.float64)
        for j in range(J):
            k, ty = float(strikes_dev[j]), int(types_dev[j])
            pay = np.where(under < k, k - under, 0.0) if ty & 1 else np.where(under > k, under - k, 0.0)
            if ty >= 2:
                pay = pay / spots
            ok = ~np.isnan(pay)
            out[3 * j], out[3 * j + 1], out[3 * j + 2] = float(pay[ok].sum()), float((pay[ok] ** 2).sum()), float(ok.sum())
        return out

    def finalize(self, sums, J, discfactor, total_paths):
        s = sums.view(J, 3)
        mean = s[:, 0] / s[:, 2]
        var = torch.clamp(s[:, 1] / s[:, 2] - mean * mean, min=0.0)
        return discfactor * mean, discfactor * torch.sqrt(var) / np.sqrt(total_paths)
