"""MMSNet: the slice of the reference's TREC-QA net that is the hot path
(examples/trec_qa_w2v_mms/do_trec_qa_clean.py:452-468, ``network_v4``):

    question ids (N,L) -> Embed --\
                                   +--> SimCross(dist_mode=2, mesure_count=mc) -> S (N,mc,L,L)
    answer   ids (N,L) -> Embed --/
      (the two Embed layers share ``w2v-weights`` / ``w2v-bias`` by param name, :461-466)

executed layer by layer the way caffe::Net does (net.cpp:535-591: ForwardFromTo /
BackwardFromTo, ClearParamDiffs :923-941, ShareWeights :944-950).  The CNN that
consumes S in the reference is out of scope, so S is treated as a loss top whose
per-element loss weights (top.diff) stand for the upstream gradient: loss = <S, dS>
(include/caffe/layer.hpp:471-479), which makes the step's scalar result well defined.
"""
import numpy as np
import torch

from .blob import Blob
from .layers import EmbedLayer, LayerParameter, SimCrossLayer


from . import _lib


class MMSNet(object):
    def __init__(self, N, L=40, D=300, mc=4, V=60002, dtype=np.float32, device="cuda", bias_term=True,
                 embed_bias=True, math=None, stage_tf32=True, deterministic=False, keep_embed_tops=True,
                 grouped_scatter=True):
        self.N, self.L, self.D, self.mc, self.V = N, L, D, mc, V
        self.dtype = np.dtype(dtype)
        self.device = torch.device(device)
        mk = lambda shape: Blob(shape, dtype=dtype, device=device)
        self.idx_q, self.idx_a = mk((N, L)), mk((N, L))
        self.q, self.a = mk(()), mk(())
        self.S = mk(())
        ep = dict(num_output=D, input_dim=V, bias_term=embed_bias,
                  weight_filler=dict(type="uniform", min=-0.08, max=0.08))
        self.embed_q = EmbedLayer(LayerParameter("Embed", name="embed_q", dtype=dtype, embed_param=ep))
        self.embed_a = EmbedLayer(LayerParameter("Embed", name="embed_a", dtype=dtype, embed_param=ep))
        self.sim = SimCrossLayer(LayerParameter(
            "SimCross", name="sim_cross", dtype=dtype, loss_weight=[1.0],
            sim_cross_param=dict(dist_mode=2, mesure_count=mc, bias_term=bias_term,
                                 weight_filler=dict(type="uniform", min=-0.1, max=0.1))))
        self.embed_q.SetUp([self.idx_q], [self.q])
        self.embed_a.SetUp([self.idx_a], [self.a])
        # param sharing by name: embed_a uses embed_q's blobs (data and diff)
        for j, b in enumerate(self.embed_q.blobs):
            self.embed_a.blobs[j].ShareData(b)
            self.embed_a.blobs[j].ShareDiff(b)
        self.sim.SetUp([self.q, self.a], [self.S])
        if math is not None:
            self.sim.set_math(math)
        # Net::ForwardBackward drives Backward right after Forward on unchanged bottoms and weights, so SimCross
        # backward may reuse the TF32-rounded operands its forward left in the workspace
        from . import _lib
        self.sim.handle.set_option(_lib.MMS_OPT_REUSE_FORWARD, 1)
        # the Embed tops go straight into SimCross (do_trec_qa_clean.py:461-468: no layer in between), so the gather
        # also writes the TF32 operand copy the contractions read (MMS_OPT_STAGE_TF32): no separate rounding pass
        if self.dtype == np.float32 and stage_tf32:
            self.embed_q.handle.set_option(_lib.MMS_OPT_STAGE_TF32, 1)
            self.embed_a.handle.set_option(_lib.MMS_OPT_STAGE_TF32, 1)
            # keep_embed_tops=False (MMS_OPT_STAGE_ONLY): q / a exist only as that operand copy; the fp32 blobs keep
            # their shape and address but are never written -- nothing in this net reads them
            if not keep_embed_tops:
                self.embed_q.handle.set_option(_lib.MMS_OPT_STAGE_ONLY, 1)
                self.embed_a.handle.set_option(_lib.MMS_OPT_STAGE_ONLY, 1)
        self.keep_embed_tops = bool(keep_embed_tops) or not (self.dtype == np.float32 and stage_tf32)
        # MMS_OPT_EMBED_DETERMINISTIC: order-independent scatter-add; the two Embed backwards then run one after the
        # other (each is the single writer of the table rows it touches)
        self.deterministic = bool(deterministic)
        if self.deterministic:
            self.embed_q.handle.set_option(_lib.MMS_OPT_EMBED_DETERMINISTIC, 1)
            self.embed_a.handle.set_option(_lib.MMS_OPT_EMBED_DETERMINISTIC, 1)
            # on the SimCross handle: dq / da rows have one writer each (no measure split with float atomics)
            self.sim.handle.set_option(_lib.MMS_OPT_EMBED_DETERMINISTIC, 1)
        # both Embed backwards as one scatter-add with the token rows grouped by id (mms_embed_backward_pair); the
        # deterministic net keeps the per-layer fixed-point kernels
        self.grouped_scatter = bool(grouped_scatter) and self.dtype == np.float32 and not self.deterministic
        self._pinned = None
        self._side = None

    # learnable params in net order, shared blobs once (net.cpp:440-530)
    def params(self):
        return list(self.embed_q.blobs) + list(self.sim.blobs)

    def set_params(self, W=None, b=None, M=None, B=None):
        if W is not None:
            self.embed_q.blobs[0].set_cpu_data(W)
        if b is not None and len(self.embed_q.blobs) > 1:
            self.embed_q.blobs[1].set_cpu_data(b)
        if M is not None:
            self.sim.blobs[0].set_cpu_data(M)
        if B is not None and len(self.sim.blobs) > 1:
            self.sim.blobs[1].set_cpu_data(B)

    # snapshots: Net::CopyTrainedLayersFrom (net.cpp:741-776) / Net::ToProto (:847-856) over .caffemodel files
    def layers(self):
        return [self.embed_q, self.embed_a, self.sim]

    def CopyTrainedLayersFrom(self, source):
        from . import formats
        return formats.copy_trained_layers_from(self.layers(), source)

    def ToProto(self, write_diff=False):
        from . import formats
        return formats.net_to_proto(self.layers(), name="mms", write_diff=write_diff)

    def Snapshot(self, path, write_diff=False):
        from . import formats
        formats.write_caffemodel(path, self.ToProto(write_diff))

    def set_upstream_gradient(self, dS):
        """top.diff of the loss top = per-element loss weights = dLoss/dS."""
        self.S.set_cpu_diff(dS)

    def set_inputs(self, idx_q, idx_a):
        self.idx_q.set_cpu_data(idx_q)
        self.idx_a.set_cpu_data(idx_a)

    def set_inputs_from_pinned(self, host_q, host_a):
        """H2D of this step's inputs from pinned host tensors (async on the current stream)."""
        self.idx_q.data.copy_(host_q, non_blocking=True)
        self.idx_a.data.copy_(host_a, non_blocking=True)

    def ClearParamDiffs(self):
        for p in self.params():
            p.diff.zero_()

    def Forward(self, with_loss=True):
        self.embed_q.Forward([self.idx_q], [self.q])
        self.embed_a.Forward([self.idx_a], [self.a])
        if with_loss:
            return self.sim.Forward([self.q, self.a], [self.S])
        saved, self.sim.loss_ = self.sim.loss_, []
        try:
            self.sim.Forward([self.q, self.a], [self.S])
        finally:
            self.sim.loss_ = saved
        return 0.0

    def Backward(self):
        self.sim.Backward([self.S], [True, True], [self.q, self.a])
        if self.grouped_scatter:
            self.embed_q.BackwardPair(self.embed_a, self.q, self.a, self.idx_q, self.idx_a)
            return
        self.embed_q.Backward([self.q], [False], [self.idx_q])
        self.embed_a.Backward([self.a], [False], [self.idx_a])

    def _prepare_weights(self, main):
        """M rounded for the coming forward on a stream of its own (it does not depend on the inputs); returns the stream
        the caller joins in front of SimCross forward."""
        if getattr(self, "_prep_stream", None) is None:
            self._prep_stream = torch.cuda.Stream()
        sp = self._prep_stream
        sp.wait_stream(main)
        with torch.cuda.stream(sp):
            self.sim.Prepare([self.q, self.a])
        return sp

    def _plan_scatter(self, side):
        """The id grouping of this step's scatter-add on ``side`` (it depends on the inputs only)."""
        if self.grouped_scatter:
            with torch.cuda.stream(side):
                self.embed_q.PlanPair(self.embed_a, self.idx_q, self.idx_a)

    def ForwardBackward(self, with_loss=True):
        loss = self.Forward(with_loss)
        self.Backward()
        return loss

    def ForwardBackwardConcurrent(self, with_loss=True, clear_diffs=True):
        r"""The same step with its independent pieces on forked streams (the reference's Net runs layers one
        after the other on the legacy stream, net.cpp:535-591; the data dependencies are all that matters):

            ClearParamDiffs, id grouping ---------------\
            round M ---\                                 +--> SimCross dq ‖ da --> Embed bwd (q, a) --\
            Embed(q) ---+--> SimCross fwd (+ loss dot) --/                     \--> SimCross dM, dB ---+--> done
            Embed(a) --/

        (SimCross backward forks its own da branch inside the library.)  Works eagerly and under stream
        capture; every side stream is joined back into the current stream before returning."""
        if self._side is None:
            self._side = (torch.cuda.Stream(), torch.cuda.Stream())
        main = torch.cuda.current_stream()
        s1, s2 = self._side
        s1.wait_stream(main)
        s2.wait_stream(main)
        sp = self._prepare_weights(main)
        if clear_diffs:
            with torch.cuda.stream(s1):
                self.ClearParamDiffs()
        self._plan_scatter(s1)
        with torch.cuda.stream(s2):
            self.embed_a.Forward([self.idx_a], [self.a])
        self.embed_q.Forward([self.idx_q], [self.q])
        main.wait_stream(s2)
        main.wait_stream(sp)
        saved, self.sim.loss_ = self.sim.loss_, []             # the loss dot (layer.hpp:471-479) goes on a side branch
        try:
            self.sim.Forward([self.q, self.a], [self.S])
        finally:
            self.sim.loss_ = saved
        loss = 0.0
        if with_loss:
            s2.wait_stream(main)
            with torch.cuda.stream(s2):
                loss = self.sim.ForwardLoss([self.S])
        main.wait_stream(s1)                                   # diffs are cleared before anything accumulates
        # dq, da first; the weight gradient (dM, dB) on a branch of its own beside the scatter-add, which needs dq / da only
        self.sim.BackwardBottoms([self.S], [self.q, self.a])
        if getattr(self, "_side3", None) is None:
            self._side3 = torch.cuda.Stream()
        s3 = self._side3
        s3.wait_stream(main)
        with torch.cuda.stream(s3):
            self.sim.BackwardParams([self.S], [self.q, self.a])
        self._embed_backward_pair(main, s2)
        main.wait_stream(s3)
        return loss

    def _embed_backward_pair(self, main, s2):
        """Both scatter-adds into the shared dW: side by side with float atomics, or -- deterministic mode -- one after
        the other, in the order MMSNet.Backward uses."""
        if self.deterministic:                                 # the order of MMSNet.Backward
            self.embed_q.Backward([self.q], [False], [self.idx_q])
            self.embed_a.Backward([self.a], [False], [self.idx_a])
            main.wait_stream(s2)                               # the loss branch still joins here
            return
        # (below the library's row threshold the pair call would run the two per-layer kernels one after the other:
        #  small batches are latency-bound, side by side is shorter)
        if self.grouped_scatter and self.idx_q.count() + self.idx_a.count() >= _lib.EMBED_GROUPED_MIN_ROWS:
            self.embed_q.BackwardPair(self.embed_a, self.q, self.a, self.idx_q, self.idx_a)
            main.wait_stream(s2)
            return
        s2.wait_stream(main)
        with torch.cuda.stream(s2):
            self.embed_a.Backward([self.a], [False], [self.idx_a])
        self.embed_q.Backward([self.q], [False], [self.idx_q])
        main.wait_stream(s2)

    def ForwardBackwardExchange(self, exch, with_loss=True, clear_diffs=True, solver=None, overlap=True):
        r"""One data-parallel step ordered by its data dependencies (the reference runs Backward layer by layer and
        only then P2PSync::on_gradients_ready, parallel.cpp:325-380): the embedding scatter-add needs dq / da only, so
        SimCross backward is split (mms_simcross_backward_bottoms / _params) and the exchange of the table gradient --
        73.5 of the 75 MB -- crosses NVLink on a private stream while dM and dB are still being computed:

            ClearParamDiffs --------------------------------\
            Embed(q) --\                                     +-> SimCross dq ‖ da --> Embed bwd (q ‖ a) --> exchange{W, b} --\
            Embed(a) ---+--> SimCross fwd (+ loss dot) -----/                     \--> SimCross dM, dB --> exchange{M, B} --+--> done

        ``exch``: GradientExchange over ``self.params()`` (peer-memory backend).  With ``solver`` (an
        AdaDeltaSolver holding the hyper-parameters) the two exchanges are the fused reduce + solver step + weight
        publication (mms_exchange_adadelta) and leave the gradients zeroed: a whole Solver::Step.  Works eagerly and
        under stream capture."""
        if self._side is None:
            self._side = (torch.cuda.Stream(), torch.cuda.Stream())
        if getattr(self, "_side3", None) is None:
            self._side3 = torch.cuda.Stream()
        main = torch.cuda.current_stream()
        s1, s2 = self._side
        s3 = self._side3
        s1.wait_stream(main)
        s2.wait_stream(main)
        sp = self._prepare_weights(main)
        if clear_diffs and solver is None:
            with torch.cuda.stream(s1):
                self.ClearParamDiffs()
        self._plan_scatter(s1)
        with torch.cuda.stream(s2):
            self.embed_a.Forward([self.idx_a], [self.a])
        self.embed_q.Forward([self.idx_q], [self.q])
        main.wait_stream(s2)
        main.wait_stream(sp)
        saved, self.sim.loss_ = self.sim.loss_, []
        try:
            self.sim.Forward([self.q, self.a], [self.S])
        finally:
            self.sim.loss_ = saved
        loss = 0.0
        if with_loss:
            s2.wait_stream(main)
            with torch.cuda.stream(s2):
                loss = self.sim.ForwardLoss([self.S])
        main.wait_stream(s1)
        self.sim.BackwardBottoms([self.S], [self.q, self.a])
        self._embed_backward_pair(main, s2)
        ne = len(self.embed_q.blobs)                           # params() = Embed blobs, then SimCross blobs
        if not overlap:                                        # the reference's order: all of Backward, then one exchange
            self.sim.BackwardParams([self.S], [self.q, self.a])
            self._exchange(exch, solver, 0, len(self.params()), channel=0)
            return loss
        s3.wait_stream(main)
        with torch.cuda.stream(s3):
            self._exchange(exch, solver, 0, ne, channel=0)
        self.sim.BackwardParams([self.S], [self.q, self.a])
        self._exchange(exch, solver, ne, len(self.params()), channel=1)
        main.wait_stream(s3)
        return loss

    @staticmethod
    def _exchange(exch, solver, first, last, channel):
        if solver is None:
            exch.allreduce(bucket=exch.bucket(first, last), channel=channel)
        else:
            exch.adadelta_step(lr_mult=solver.lr_mult, decay_mult=solver.decay_mult, base_lr=solver.GetLearningRate(),
                               momentum=solver.momentum, delta=solver.delta, weight_decay=solver.weight_decay,
                               iter_size=solver.iter_size, bucket_blobs=(first, last), channel=channel, clear_diffs=True)

    def capture_exchange_step(self, exch, with_loss=True, clear_diffs=True, solver=None, host_inputs=None, overlap=True,
                              prefetch_inputs=False):
        """Records ForwardBackwardExchange as ONE CUDA graph (the exchange kernels keep their epoch in device memory,
        so a replay is a new exchange).  Two eager passes first: they size the workspaces, fill the tensor-map cache
        and, with ``solver``, allocate the history -- and they are REAL steps, so every rank must call this the same
        number of times.  ``host_inputs=(host_q, host_a)`` adds the H2D copies of the pinned id tensors and the D2H
        copy of the loss to the graph (replay_exchange_from_host)."""
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        self.sim.defer_loss_ = True
        try:
            with torch.cuda.stream(side):
                for _ in range(2):
                    self.ForwardBackwardExchange(exch, with_loss, clear_diffs, solver, overlap)
                exch.check()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                ins = self._inputs_into_graph(host_inputs, prefetch_inputs) if host_inputs is not None else None
                self.ForwardBackwardExchange(exch, with_loss, clear_diffs, solver, overlap)
                self._inputs_handover(ins)
                if host_inputs is not None and with_loss:
                    if getattr(self, "_host_loss", None) is None:
                        self._host_loss = torch.zeros(1, dtype=torch.float32).pin_memory()
                    self._host_loss.copy_(self.sim.loss_dev_[0], non_blocking=True)
        finally:
            self.sim.defer_loss_ = False
        return graph

    # -- CUDA-graph replay of the whole step ----------------------------------------------
    # The step is ~15 short kernels; issued one by one from the host it is launch-bound
    # (the reference has the same problem in the small: 2*N*mc host BLAS calls).  Recording
    # ClearParamDiffs + Forward + Backward once and replaying the graph removes the host from
    # the loop.  Blob storage must not be re-allocated between capture and replay.
    def capture_train_step(self, solver, with_loss=True):
        """Records a whole solver iteration: Forward + Backward + ``solver.ApplyUpdate(clear_diffs=True)`` -- the
        fused optimizer launch also zeroes every diff for the next iteration, so the step contains no separate
        ClearParamDiffs pass (Solver::Step, solver.cpp:195-260).  Diffs must be zero when the first replay starts."""
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        self.sim.defer_loss_ = True
        try:
            with torch.cuda.stream(side):
                # warm-up without the optimizer (sizes the workspaces, fills the tensor-map cache): weights, history
                # and the iteration count are exactly what they were when the first replay runs
                for _ in range(2):
                    self.ForwardBackwardConcurrent(with_loss, clear_diffs=False)
                    self.ClearParamDiffs()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            it = solver.iter
            with torch.cuda.graph(graph, stream=side):
                self.ForwardBackwardConcurrent(with_loss, clear_diffs=False)
                solver.ApplyUpdate(clear_diffs=True)
            solver.iter = it                                # recording is not an iteration
        finally:
            self.sim.defer_loss_ = False
        self._graph_train = graph
        return graph

    def replay_train_step(self, solver=None):
        self._graph_train.replay()
        if solver is not None:
            solver.iter += 1

    def _inputs_into_graph(self, host_inputs, prefetch):
        """Called inside a capture, before the step: the H2D copies of the pinned id tensors.  ``prefetch``: they go into
        staging buffers on a side stream BESIDE the step's kernels (the ids are the first thing a step needs, so a copy in
        front of it is fully exposed) and `_inputs_handover` moves them into the id blobs at the end of the step -- every
        replay computes on the ids its predecessor fetched and fetches the next ones (input double buffering: the caller
        refills the pinned tensors with batch k+1 before replay k; prime with set_inputs)."""
        host_q, host_a = host_inputs
        assert host_q.is_pinned() and host_a.is_pinned()
        if not prefetch:
            self.set_inputs_from_pinned(host_q, host_a)
            return None
        if getattr(self, "_in_stage", None) is None:
            self._in_stage = (torch.empty_like(self.idx_q.data), torch.empty_like(self.idx_a.data))
            self._in_stream = torch.cuda.Stream()
        self._in_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._in_stream):
            self._in_stage[0].copy_(host_q.view_as(self._in_stage[0]), non_blocking=True)
            self._in_stage[1].copy_(host_a.view_as(self._in_stage[1]), non_blocking=True)
        return self._in_stream

    def _inputs_handover(self, in_stream):
        if in_stream is not None:
            torch.cuda.current_stream().wait_stream(in_stream)
            self.idx_q.data.copy_(self._in_stage[0])
            self.idx_a.data.copy_(self._in_stage[1])

    def capture(self, with_loss=True, clear_diffs=True, host_inputs=None, prefetch_inputs=False):
        """Records one step.  With `host_inputs=(host_q, host_a)` (pinned tensors) a second graph is recorded
        that also contains the H2D copies of the two id tensors and the D2H copy of the loss scalar into a pinned
        buffer, so that an end-to-end step is ONE graph launch and one stream synchronisation
        (`replay_from_host`).  `prefetch_inputs`: see `_inputs_into_graph`."""
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        self.sim.defer_loss_ = True
        try:
            with torch.cuda.stream(side):
                for _ in range(2):                      # sizes scratch, fills the tensor-map cache
                    self.ForwardBackwardConcurrent(with_loss, clear_diffs)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                self.ForwardBackwardConcurrent(with_loss, clear_diffs)
        finally:
            self.sim.defer_loss_ = False
        self._graph = graph
        self._graph_with_loss = with_loss
        self._graph_host = None
        if host_inputs is not None:
            host_q, host_a = host_inputs
            assert host_q.is_pinned() and host_a.is_pinned()
            self._host_loss = torch.zeros(1, dtype=torch.float32).pin_memory()
            self.sim.defer_loss_ = True
            try:
                g2 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g2, stream=side):
                    ins = self._inputs_into_graph(host_inputs, prefetch_inputs)
                    self.ForwardBackwardConcurrent(with_loss, clear_diffs)
                    self._inputs_handover(ins)
                    if with_loss:
                        self._host_loss.copy_(self.sim.loss_dev_[0], non_blocking=True)
            finally:
                self.sim.defer_loss_ = False
            self._graph_host = g2
            self._graph_stream = side
        return graph

    def replay_from_host(self):
        """One recorded end-to-end step: H2D of the pinned inputs given to `capture`, the step, D2H of the loss."""
        self._graph_host.replay()
        torch.cuda.current_stream().synchronize()
        return float(self._host_loss[0]) if self._graph_with_loss else 0.0

    def replay(self, read_loss=True):
        """One recorded step.  Returns the loss (D2H of one scalar + sync) when asked to."""
        self._graph.replay()
        if read_loss and self._graph_with_loss:
            return float(self.sim.loss_dev_[0].item())
        return 0.0
