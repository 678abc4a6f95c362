"""Generate tests/golden/solve_reference.npz by EXECUTING the reference's own PCG driver,
``solve()`` of ``/root/reference/src/toast/ops/mapmaker_solve.py:524-755``.

``import toast`` is impossible here, but ``solve()`` only needs numpy, a logger / timer, and
duck-typed ``data`` / ``lhs_op`` / AmplitudesMap objects.  Its definition is lifted out of the
reference source with ``ast`` and executed where it lies (nothing is copied into this repository)
against stand-ins whose LHS operator and preconditioner call the reference's COMPILED kernels
(oracle/_ref) in the order SolverLHS does.  Every ``dot`` is recorded at full precision (the
function itself only logs six digits), which gives the residual history.  Run in the build
container only:

    python tests/golden/make_golden_solve.py
"""

import ast
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
os.environ.setdefault("OMP_NUM_THREADS", "1")
REF_SRC = "/root/reference/src/toast/ops/mapmaker_solve.py"

DOTS = []  # (is_self_dot, value) in call order


class Amp:
    """One template's amplitudes (templates/amplitudes.py:201-274, 523-571): local values and
    flags; dot() skips flagged entries."""

    def __init__(self, local, flags):
        self.local = local
        self.local_flags = flags
        self.n_local = self.n_global = len(local)


class AmpMap(dict):
    """templates/amplitudes.py:956-975 (AmplitudesMap): a dict of Amp with vector arithmetic."""

    def duplicate(self):
        out = AmpMap()
        for k, v in self.items():
            out[k] = Amp(v.local.copy(), v.local_flags.copy())
        return out

    def reset(self):
        for v in self.values():
            v.local[:] = 0

    def dot(self, other):
        total = 0.0
        for k, v in self.items():
            good = v.local_flags == 0
            total += float(np.dot(np.where(good, v.local, 0), np.where(good, other[k].local, 0)))
        DOTS.append((other is self, total))
        return total

    def __iadd__(self, other):
        for k, v in self.items():
            v.local += other[k].local
        return self

    def __isub__(self, other):
        for k, v in self.items():
            v.local -= other[k].local
        return self

    def __imul__(self, other):
        for v in self.values():
            v.local *= other
        return self


class _Quiet:
    def __getattr__(self, name):
        return lambda *a, **k: None


class _Logger:
    @staticmethod
    def get():
        return _Quiet()


class _Timer(_Quiet):
    pass


def load_reference_solve():
    tree = ast.parse(open(REF_SRC).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "solve")
    fn.decorator_list = []
    ns = {"np": np, "Logger": _Logger, "Timer": _Timer, "AmplitudesMap": AmpMap}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), REF_SRC, "exec"), ns)
    return ns["solve"]


class _Data(dict):
    class comm:  # noqa: N801
        comm_world = None
        world_rank = 0


def run_reference_solve(pb, K, O, rhs, n_iter_max, n_iter_min=3, convergence=1.0e-12):
    """Returns (amplitudes, residual history) of the reference's solve() on problem ``pb``."""
    solve = load_reference_solve()

    class TemplateMatrix:
        amplitudes = None

        @staticmethod
        def apply_precond(a_in, a_out):
            K.template_offset_apply_diag_precond(pb.offset_var, a_in["baselines"].local,
                                                 a_in["baselines"].local_flags,
                                                 a_out["baselines"].local, False)

    class LhsOp:
        name = "lhs"
        out = None
        template_matrix = TemplateMatrix

        @staticmethod
        def apply(data, detectors=None):
            a = data[TemplateMatrix.amplitudes]["baselines"].local
            data[LhsOp.out]["baselines"].local[:] = O.solver_lhs(pb, K, a)

    data = _Data()
    data["rhs"] = AmpMap(baselines=Amp(rhs.copy(), pb.amp_flags.copy()))
    del DOTS[:]
    solve(data, None, LhsOp, "rhs", "result", convergence=convergence, n_iter_max=n_iter_max,
          n_iter_min=n_iter_min)
    sq0 = DOTS[0][1]
    hist = [v / sq0 for self_dot, v in DOTS[2:] if self_dot]
    return data["result"]["baselines"].local.copy(), np.array(hist)


def main():
    from oracle import toast_oracle as O
    from toast_b200 import synthetic as S

    ref = O.load_ref()
    assert ref is not None, "build oracle/_ref first"
    out = {}
    for tag, (name, n_det, n_samp, nside, n_iter) in dict(
            c1=("c1", 4, 6000, 64, 25), c2=("c2", 4, 12000, 64, 12)).items():
        obs = S.make_observation(name, n_det=n_det, n_samp=n_samp, nside=nside, eps_max=0.03)
        pb = O.build_problem(obs, ref)
        rhs = O.solver_rhs(pb, ref, obs["signal"])
        amps, hist = run_reference_solve(pb, ref, O, rhs, n_iter)
        out[f"{tag}_args"] = np.array([n_det, n_samp, nside, n_iter])
        out[f"{tag}_rhs"] = rhs
        out[f"{tag}_amplitudes"] = amps
        out[f"{tag}_history"] = hist
        print(tag, len(hist), "iterations, final relative residual", hist[-1])
    np.savez_compressed(os.path.join(HERE, "solve_reference.npz"), **out)


if __name__ == "__main__":
    main()
