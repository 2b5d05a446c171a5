"""mms_answer_selection_b200 -- B200-native (sm_100a) implementation of the MMS
answer-selection hot path of lxmeng/mms_answer_selection (a Caffe fork):
Embed -> SimCross / SimMatrix bilinear scoring -> PairRankLoss / FM, forward and
backward, plus data-parallel gradient exchange and candidate scoring.

The compute lives in ``libmms_b200.so`` (hand-written CUDA behind the C-ABI declared in
``include/mms_b200.h``).  This package is the host-side mirror of the reference's Caffe
Layer interface (``LayerSetUp / Reshape / Forward / Backward``, Blob data/diff) used by
the tests, the benchmark and ``torch.distributed`` data-parallel runs; PyTorch is used
for device memory, streams and NCCL only.  There is no CPU fallback: without the built
library or without a B200 the layers raise.
"""
from ._lib import MMSError, lib, lib_path  # noqa: F401
from .blob import Blob  # noqa: F401
from .layers import (AUCLayer, BNLayer, ConvolutionLayer, DropoutLayer, EmbedLayer, PoolingLayer, TanHLayer, FMLayer, Layer, LayerParameter, MAPLayer, MRRLayer,  # noqa: F401
                     PairRankLossLayer, RankAccuracyLayer, SimCrossLayer, SimMatrixLayer, create_layer)
from . import parallel  # noqa: F401
from .net import MMSNet  # noqa: F401
from .parallel import GradientExchange  # noqa: F401
from .sentnet import SentenceVectorNet  # noqa: F401
from .solver import AdaDeltaSolver  # noqa: F401

__all__ = ["MMSError", "lib", "lib_path", "Blob", "Layer", "LayerParameter", "EmbedLayer",
           "SimCrossLayer", "SimMatrixLayer", "PairRankLossLayer", "FMLayer", "create_layer", "MMSNet", "AdaDeltaSolver", "GradientExchange",
           "MAPLayer", "MRRLayer", "AUCLayer", "RankAccuracyLayer",
           "ConvolutionLayer", "DropoutLayer", "BNLayer", "PoolingLayer", "TanHLayer", "SentenceVectorNet"]
