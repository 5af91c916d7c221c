"""Local Richardson extrapolation over any integrator of this package (reference:
/root/reference/desolver/integrators/integrator_types.py:420-613, integrator_template.py:110-115).

`generate_richardson_integrator(basis, richardson_iter)` returns a class that advances one step of size h by running the
BASIS integrator over 1, 2, 4, ... subdivisions of h and extrapolating the increments (Aitken-Neville table,
`stage_values[m, n] = s[m, n-1] + (s[m, n-1] - s[m-1, n-1]) / (2^n - 1)`); the difference of the last two diagonal
entries is the error estimate fed to the step-size controller (`update_timestep`).  Like in the reference this is a
wrapper around the integrator PROTOCOL: every sub-step is one call of the basis integrator -- i.e. the stage-level
kernels of libdesolver_b200.so with the right-hand side in between -- and only the extrapolation table itself (a few
axpy's over the state) is host-side tensor arithmetic.  `OdeSystem` drives such an integrator through its protocol loop
(`OdeSystem._run_protocol`): single systems, one host visit per step.
"""
import abc

from .integrator_types import IntegratorTemplate

__all__ = ["RichardsonIntegratorTemplate", "generate_richardson_integrator"]


class RichardsonIntegratorTemplate(IntegratorTemplate, abc.ABC):
    symplectic = False

    @property
    def adaptive(self):
        return True


def generate_richardson_integrator(basis_integrator, richardson_iter=2):
    import torch

    class RichardsonExtrapolatedIntegrator(RichardsonIntegratorTemplate):
        __alt_names__ = ("Local Richardson Extrapolation of {}".format(basis_integrator.__name__),)
        symplectic = basis_integrator.symplectic

        def __init__(self, sys_dim, **kwargs):
            super().__init__()
            from .. import backend as D
            self.dim = tuple(int(i) for i in sys_dim)
            self.numel = 1
            for i in self.dim:
                self.numel *= int(i)
            self.dtype = kwargs.get("dtype")
            self.device = kwargs.get("device", None)
            eps_dtype = self.dtype if self.dtype is not None else "float64"
            self.rtol = float(kwargs["rtol"]) if kwargs.get("rtol") is not None else 32 * D.epsilon(eps_dtype)
            self.atol = float(kwargs["atol"]) if kwargs.get("atol") is not None else 32 * D.epsilon(eps_dtype)
            self.richardson_iter = int(richardson_iter)
            basis_kwargs = {k: v for k, v in kwargs.items() if k != "staggered_mask" or basis_integrator.symplectic}
            self.basis_integrators = [basis_integrator(sys_dim, **basis_kwargs) for _ in range(self.richardson_iter)]
            self.basis_order = self.basis_integrators[0].order
            if basis_integrator.symplectic:
                self.staggered_mask = self.basis_integrators[0].staggered_mask
            self.dState = None
            self.dTime = None
            self.initial_state = self.initial_rhs = self.initial_time = self.final_rhs = None
            self._adaptive = True
            self._interpolants = None
            self._interpolant_times = None
            self.solver_dict = dict(safety_factor=0.5 if self.basis_integrators[0].is_implicit else 0.9, atol=self.atol,
                                    rtol=self.rtol, order=self.basis_order + self.richardson_iter // 2)

        def dense_output(self):
            return self._interpolant_times, self._interpolants

        def check_converged(self, initial_state, diff, prev_error):
            err_estimate = torch.max(torch.abs(diff))
            relerr = torch.max(self.atol + self.rtol * torch.abs(initial_state))
            if prev_error is None or (bool(err_estimate > relerr) and bool(err_estimate <= torch.max(torch.abs(prev_error)))):
                return diff, False
            return diff, True

        def subdiv_step(self, int_num, rhs, initial_time, initial_state, timestep, constants, num_intervals):
            dt_now, dstate_now = 0.0, 0.0
            dtstep = timestep / num_intervals
            self._interpolants, self._interpolant_times = [], []
            basis = self.basis_integrators[int_num]
            for _ in range(num_intervals):
                _dt, (dt_z, dy_z) = basis(rhs, initial_time + dt_now, initial_state + dstate_now, constants, dtstep)
                dt_now = dt_now + dt_z
                dstate_now = dstate_now + dy_z
                t_i, interp = basis.dense_output()
                self._interpolant_times.append(t_i)
                self._interpolants.append(interp)
            return dtstep, (dt_now, dstate_now)

        def adaptive_richardson(self, rhs, t, y, constants, timestep):
            R = self.richardson_iter
            table = [[None] * R for _ in range(R)]
            _dt0, (dt_z, dy_z) = self.subdiv_step(0, rhs, t, y, timestep, constants, 1)
            if bool(dt_z < timestep):
                timestep = dt_z
            table[0][0] = dy_z
            prev_error = None
            m = n = 0
            for m in range(1, R):
                table[m][0] = self.subdiv_step(m, rhs, t, y, timestep, constants, 1 << m)[1][1]
                for n in range(1, m + 1):
                    table[m][n] = table[m][n - 1] + (table[m][n - 1] - table[m - 1][n - 1]) / ((1 << n) - 1)
                self.solver_dict["order"] = self.basis_order + m + 1
                if m >= 3:
                    prev_error, t_conv = self.check_converged(table[m][n], table[m - 1][m - 1] - table[m][m], prev_error)
                    if t_conv:
                        break
            self.solver_dict["num_richardson_iterations"] = m
            return timestep, (timestep, table[m - 1][n - 1]), table[m - 1][m - 1] - table[m][m]

        def __call__(self, rhs, initial_time, initial_state, constants, timestep):
            f64 = dict(dtype=torch.float64, device=initial_state.device if hasattr(initial_state, "device") else self.device)
            initial_state = torch.as_tensor(initial_state, **f64)
            initial_time = torch.as_tensor(initial_time, **f64)
            timestep = torch.as_tensor(timestep, **f64)
            dt0, (dt_z, dy_z), diff = self.adaptive_richardson(rhs, initial_time, initial_state, constants, timestep)
            self.dState = dy_z + 0.0
            self.dTime = dt_z.clone()
            self.initial_state, self.initial_time = initial_state, initial_time
            self.solver_dict.update(diff=diff, initial_state=initial_state, initial_time=initial_time, timestep=self.dTime,
                                    dState=self.dState)
            # (like the reference, the wrapper never wipes the controller history: scaling EMA and the three-term product
            # run over the whole integration, integrator_types.py:543-552)
            new_timestep, redo_step = self.update_timestep()
            if self.symplectic:
                timestep = dt0
                next_timestep = dt0.clone()
                if bool((0.8 * new_timestep + 0.2 * timestep) < next_timestep):
                    while bool(new_timestep < next_timestep):
                        next_timestep = next_timestep / 2.0
                else:
                    while bool((0.8 * new_timestep + 0.2 * timestep) > 2 * next_timestep):
                        next_timestep = next_timestep * 2.0
                    redo_step = False
            else:
                next_timestep = new_timestep
            if redo_step:
                timestep, (self.dTime, self.dState) = self(rhs, initial_time, initial_state, constants, next_timestep)
            else:
                timestep = next_timestep
            return timestep, (self.dTime, self.dState)

        @property
        def order(self):
            return self.basis_order + self.richardson_iter // 2

        @property
        def is_implicit(self):
            return self.basis_integrators[0].is_implicit

        @property
        def is_explicit(self):
            return self.basis_integrators[0].is_explicit

        @property
        def is_adaptive(self):
            return True

        @property
        def stages(self):
            return self.basis_integrators[0].stages * self.richardson_iter

    RichardsonExtrapolatedIntegrator.__name__ = RichardsonExtrapolatedIntegrator.__qualname__ = \
        "RichardsonExtrapolated_{}_Integrator".format(basis_integrator.__name__)
    return RichardsonExtrapolatedIntegrator
