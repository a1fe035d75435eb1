"""SciPy's L-BFGS-B as a reverse-communication state machine.

`scipy.optimize.minimize(method='L-BFGS-B')` -- what the reference's `gpflow.optimizers.Scipy().minimize` runs
(lcgp.py:537-540) -- drives the routine `setulb`, which RETURNS to its caller whenever it needs the objective and
gradient at a new point.  `LbfgsbMachine` wraps one such instance so that many independent optimizations can be
advanced in lock-step by one host thread, all their pending evaluations served by ONE batched device call
(lcgp_b200.batched.fit_emulators).  The iterates are those of `scipy.optimize.minimize` bit for bit (same routine,
same defaults, same stopping tests; tests/test_host_logic.py::test_lbfgsb_machine_equals_scipy_minimize).
"""
from __future__ import annotations

import numpy as np
from scipy.optimize import _lbfgsb          # the compiled routine SciPy's own driver calls (scipy >= 1.15 layout)


class LbfgsbMachine:
    """One unbounded L-BFGS-B minimisation.  Protocol:
        m = LbfgsbMachine(x0)
        while m.advance():            # True: the objective and gradient at m.x are wanted
            m.supply(f(m.x), g(m.x))
        m.x, m.f, m.nit, m.nfev, m.success, m.message
    Options and defaults as scipy.optimize.minimize(method='L-BFGS-B')."""

    def __init__(self, x0, maxcor=10, ftol=2.2204460492503131e-09, gtol=1e-5, maxfun=15000, maxiter=15000, maxls=20):
        x0 = np.asarray(x0, dtype=np.float64).ravel()
        n = x0.size
        self.m, self.maxfun, self.maxiter, self.maxls = int(maxcor), int(maxfun), int(maxiter), int(maxls)
        self.factr = ftol / np.finfo(float).eps
        self.pgtol = gtol
        try:                                    # integer width of the compiled routine (as SciPy's own driver picks it)
            from scipy.optimize._lbfgsb_py import HAS_ILP64
        except ImportError:
            HAS_ILP64 = False
        idt = np.int64 if HAS_ILP64 else np.int32
        self.x = np.array(x0, dtype=np.float64)
        self.f = np.array(0.0, dtype=np.float64)
        self.g = np.zeros(n, dtype=np.float64)
        self.nbd = np.zeros(n, dtype=idt)
        self.low = np.zeros(n, dtype=np.float64)
        self.up = np.zeros(n, dtype=np.float64)
        m = self.m
        self.wa = np.zeros(2 * m * n + 5 * n + 11 * m * m + 8 * m, np.float64)
        self.iwa = np.zeros(3 * n, dtype=idt)
        self.task = np.zeros(2, dtype=idt)
        self.ln_task = np.zeros(2, dtype=idt)
        self.lsave = np.zeros(4, dtype=idt)
        self.isave = np.zeros(44, dtype=idt)
        self.dsave = np.zeros(29, dtype=np.float64)
        self.nit = 0
        self.nfev = 0
        self.done = False
        self.success = False
        self.message = ''

    def advance(self) -> bool:
        """Runs the routine until it wants f, g at self.x (-> True) or stops (-> False)."""
        if self.done:
            return False
        while True:
            _lbfgsb.setulb(self.m, self.x, self.low, self.up, self.nbd, self.f, self.g, self.factr, self.pgtol, self.wa,
                           self.iwa, self.task, self.lsave, self.isave, self.dsave, self.maxls, self.ln_task)
            if self.task[0] == 3:                    # FG: evaluate at self.x
                return True
            if self.task[0] == 1:                    # NEW_X: an iteration was completed
                self.nit += 1
                if self.nit >= self.maxiter:
                    self.task[0], self.task[1] = 5, 504
                elif self.nfev > self.maxfun:
                    self.task[0], self.task[1] = 5, 502
                continue
            break
        self.done = True
        self.success = bool(self.task[0] == 4)       # CONVERGENCE
        self.message = f'task {int(self.task[0])}/{int(self.task[1])}'
        return False

    def supply(self, f, g):
        """Objective and gradient at self.x (answer to advance() == True)."""
        self.f = np.array(float(f), dtype=np.float64)
        self.g = np.asarray(g, dtype=np.float64).copy()
        self.nfev += 1
