"""Example harness with the reference's interface (docs/call_model.py:5-86 = illustration-examples/call_model.py):
`LCGPRun(runno=..., data=dict(xtrain, ytrain, xtest, ytest[, ytrue, ystd]), submethod, robust, err_struct, num_latent,
var_threshold).define_model() / .train() / .predict()`.  Outputs of predict are numpy arrays, (p, n0) -- or their
transposes with as_pxn=True, exactly as the reference's wrapper returns them."""
from __future__ import annotations

from .model import LCGP


class SuperRun:
    """Holds one train / test split and the fitted model (docs/call_model.py:5-32)."""

    def __init__(self, runno: str, data, verbose=False, **kwargs):
        self.data = data
        self.xtrain, self.ytrain = data['xtrain'], data['ytrain']
        self.xtest, self.ytest = data['xtest'], data['ytest']
        for opt in ('ytrue', 'ystd'):
            if opt in data:
                setattr(self, opt, data[opt])
        self.runno = runno
        self.model = None
        self.modelname = ''
        self.n = self.xtrain.shape[0]
        self.num_output = self.ytrain.shape[0]
        self.verbose = verbose

    def define_model(self):
        pass

    def train(self):
        pass

    def predict(self):
        pass


class LCGPRun(SuperRun):
    """docs/call_model.py:35-86.  Extra keyword arguments are swallowed like in the reference (SURVEY B-10: the
    illustrations pass diag_error_structure= / robust_mean=, which therefore never reach the model)."""

    def __init__(self, submethod='full', robust=True, err_struct=None, num_latent=None, var_threshold=None, **kwargs):
        super().__init__(**kwargs)
        self.modelname = 'LCGP' + ('_robust' if robust else '')
        self.num_latent = num_latent
        self.var_threshold = var_threshold
        self.submethod = submethod
        self.robust = robust
        self.err_struct = err_struct

    def define_model(self):
        self.model = LCGP(y=self.ytrain, x=self.xtrain, parameter_clamp_flag=False, q=self.num_latent,
                          var_threshold=self.var_threshold, diag_error_structure=self.err_struct,
                          robust_mean=self.robust, submethod=self.submethod)

    def train(self):
        self.model.fit(verbose=self.verbose)

    def predict(self, train: bool = False, return_fullcov: bool = False, as_pxn: bool = False):
        out = self.model.predict(self.xtrain if train else self.xtest, return_fullcov=return_fullcov)
        moments = [t.numpy() for t in out[:3]]
        if as_pxn:
            moments = [a.T for a in moments]
        if return_fullcov:
            full = out[3]
            return (*moments, full.numpy() if full is not None else None)
        return tuple(moments)
