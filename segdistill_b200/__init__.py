"""segdistill_b200 - B200-native (sm_100a) kernels for SegDistill's dense distillation losses.

Only the hot path lives here: the CUDA kernels + C ABI (``csrc/``, ``include/segdistill.h``),
the ctypes binding (``_cabi``), the autograd bridges (``functional``) and the host-side mirror of
the reference's loss-module / dispatcher interface (``losses``, ``opts``, ``dist``), and the
student head's supervised loss next to it (``seg_losses``: resize + cross-entropy + accuracy).
"""
from .losses import (ATLoss, CDLoss, CDMSELoss, CGDCorrLoss, CGDLoss, CGDLossWS, FeatureMSELoss, IFVDLoss,  # noqa: F401
                     KLDLoss, PDLoss)
from .opts import (DistillationLoss, DistillationLossMT, Extractor, ExtractorMT, LOSS_CLASSES,  # noqa: F401
                   build_criterion)
from .seg_losses import CrossEntropyLoss, decode_head_losses, seg_ce_loss  # noqa: F401

__version__ = '0.1.0'
