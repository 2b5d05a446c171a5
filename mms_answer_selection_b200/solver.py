"""AdaDeltaSolver: the host-side mirror of the reference's solver step for the learnable blobs of the MMS path
(src/caffe/solvers/sgd_solver.cpp:102-116 ``SGDSolver::ApplyUpdate``, src/caffe/solvers/adadelta_solver.cpp:25-107
``AdaDeltaSolver::ComputeUpdateValue``, hyper-parameters of examples/trec_qa_w2v_mms/do_trec_qa_clean.py:48-59:
base_lr 1.0, momentum 0.95, delta 5e-7, weight_decay 5e-4, lr_policy "fixed").

The reference walks every blob about twenty times per iteration (Normalize, Regularize, eight math passes of the
update, Net::Update, and Net::ClearParamDiffs at the top of the next iteration, solver.cpp:203).  Here one launch of
``mms_adadelta_step_*`` per blob does all of it -- including the 1/solver_count scaling that P2PSync applies to the
summed gradient (parallel.cpp:377) when the step follows a sum all-reduce.  The two history blobs per parameter are
allocated like ``SGDSolver::PreSolve`` / ``AdaDeltaSolver::AdaDeltaPreSolve`` do (history_[i] and
history_[i + n])."""
import ctypes

import torch

from ._lib import Handle, c_p, check, lib


class AdaDeltaSolver(object):
    def __init__(self, params, lr_mult=None, decay_mult=None, base_lr=1.0, momentum=0.95, delta=5e-7,
                 weight_decay=5e-4, iter_size=1):
        """``params``: learnable Blobs (shared blobs once, net order).  ``lr_mult`` / ``decay_mult``: per-blob
        multipliers (ParamSpec, caffe.proto:283-305), default 1."""
        self.params = list(params)
        n = len(self.params)
        self.lr_mult = list(lr_mult) if lr_mult is not None else [1.0] * n
        self.decay_mult = list(decay_mult) if decay_mult is not None else [1.0] * n
        if len(self.lr_mult) != n or len(self.decay_mult) != n:
            raise ValueError("one lr_mult / decay_mult per learnable blob")
        self.base_lr, self.momentum, self.delta = float(base_lr), float(momentum), float(delta)
        self.weight_decay, self.iter_size = float(weight_decay), int(iter_size)
        self.iter = 0
        # history_[i]: gradient history, history_[n + i]: update history (adadelta_solver.cpp:12-22)
        self.history = [torch.zeros_like(b.data) for b in self.params] + [torch.zeros_like(b.data) for b in self.params]
        self.handle = Handle()

    def GetLearningRate(self):
        return self.base_lr                      # lr_policy "fixed" (sgd_solver.cpp:27-30)

    def ApplyUpdate(self, grad_scale=1.0, clear_diffs=False):
        """One solver iteration's update of every learnable blob.  ``grad_scale`` multiplies the gradients first
        (1/world after a sum all-reduce; the 1/iter_size of Normalize is applied on top); ``clear_diffs`` leaves the
        diffs zeroed for the next iteration instead of holding the applied update."""
        rate = self.GetLearningRate()
        scale = float(grad_scale) / self.iter_size
        self.handle.set_stream(torch.cuda.current_stream().cuda_stream)
        L = lib()
        n = len(self.params)
        for i, b in enumerate(self.params):
            if not b.data.is_cuda:
                raise RuntimeError("the solver step runs on the GPU only (no CPU fallback)")
            f32 = b.data.dtype == torch.float32
            fn = L.mms_adadelta_step_f32 if f32 else L.mms_adadelta_step_f64
            real = ctypes.c_float if f32 else ctypes.c_double
            check(fn(self.handle.ptr, c_p(b.data.data_ptr()), c_p(b.diff.data_ptr()),
                     c_p(self.history[i].data_ptr()), c_p(self.history[n + i].data_ptr()), b.count(),
                     real(scale), real(self.weight_decay * self.decay_mult[i]), real(self.momentum),
                     real(self.delta), real(rate * self.lr_mult[i]), 1 if clear_diffs else 0))
        self.iter += 1
