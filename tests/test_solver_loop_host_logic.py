"""`Destriper.solve` -- the PCG loop of mapmaker_solve.py:524-755 as the device solver drives it,
with the next direction and the next LHS enqueued speculatively while the host reads r.r back --
in the CPU suite: the five device steps (LHS, dot, x / r / s update, preconditioner, direction)
are stood in for by the oracle on CPU tensors, the LOOP is the product's.  Its residual history
and amplitudes must be the oracle's own `solve` bit for bit, including the early exits
(convergence, the stall test every ten iterations) after which the speculative work must have
left x untouched."""

import numpy as np
import pytest
import torch

from helpers import O, S
from toast_b200.solver import Destriper


class _Event:
    def record(self, stream=None):
        pass

    def synchronize(self):
        pass


class _OracleStepDestriper(Destriper):
    """solver.Destriper with every device step replaced by the oracle (no native handle)."""

    def __init__(self, pb):
        self.pb = pb
        self.device = torch.device("cpu")
        self.n_amp = pb.n_amp
        self.world, self.pipeline, self.prior = 1, False, None
        self.amp_flags = torch.from_numpy(pb.amp_flags)
        self.lhs_calls = 0

    def lhs(self, amps_in, amps_out, timers=None):
        self.lhs_calls += 1
        amps_out.copy_(torch.from_numpy(O.solver_lhs(self.pb, O, amps_in.numpy().copy())))
        return amps_out

    def dot(self, a, b, out):
        out[0] = float(O.amp_dot(a.numpy(), b.numpy(), self.pb.amp_flags))

    def precond(self, r, s):
        O.template_offset_apply_diag_precond(self.pb.offset_var, r.numpy(), self.pb.amp_flags,
                                             s.numpy(), False)

    def update(self, st):
        # tb_pcg_update: alpha = delta / d.q; x += alpha d; r -= alpha q; s = M^-1 r; (r.r, s.r)
        alpha = float(st.delta[0]) / float(st.dq[0])
        st.x.add_(st.d * alpha)
        st.r.sub_(st.q * alpha)
        self.precond(st.r, st.s)
        self.dot(st.r, st.r, st.sums[0:1])
        self.dot(st.s, st.r, st.sums[1:2])

    def advance_direction(self, st):
        beta = float(st.sums[1]) / float(st.delta[0])
        st.d.mul_(beta).add_(st.s)
        st.delta.copy_(st.sums[1:2])


@pytest.fixture
def cpu_solver(monkeypatch):
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self, raising=False)
    monkeypatch.setattr(torch.cuda, "Event", lambda *a, **k: _Event())


@pytest.mark.parametrize("n_iter_max,convergence,n_iter_min",
                         [(6, 1e-30, 3), (40, 1e-8, 3), (35, 1e-30, 3), (0, 1e-12, 3)])
def test_solve_loop_matches_the_reference_loop(cpu_solver, n_iter_max, convergence, n_iter_min):
    obs = S.make_observation("c1", n_det=4, n_samp=6000, nside=32, eps_max=0.03)
    pb = O.build_problem(obs, O)
    rhs = O.solver_rhs(pb, O, obs["signal"])
    amps_ref, hist_ref = O.solve(pb, O, rhs, convergence=convergence, n_iter_max=n_iter_max,
                                 n_iter_min=n_iter_min)
    ds = _OracleStepDestriper(pb)
    amps, hist = ds.solve(torch.from_numpy(rhs.copy()), convergence=convergence,
                          n_iter_max=n_iter_max, n_iter_min=n_iter_min)
    assert hist == hist_ref
    np.testing.assert_array_equal(amps.numpy(), amps_ref)
    # one LHS for the starting residual, one per iteration, and at most one speculative extra
    assert len(hist) + 1 <= ds.lhs_calls <= len(hist) + 2
    if 0 < len(hist) < n_iter_max:
        assert ds.lhs_calls == len(hist) + 2     # an early exit leaves the speculative LHS behind
