"""vtc_b200 -- B200-native contrastive-retrieval hot path of unitaryai/VTC.

Drop-in modules (same symbols and signatures as the reference):
    vtc_b200.model.loss.clip_loss                         <- model/loss.py:18-22
    vtc_b200.model.metric.RecallAtK                       <- model/metric.py:103-187
    vtc_b200.model.model.{normalize, PretrainedCLIPBase, PretrainedCLIP, PretrainedCLIP_finaltf}
                                                          <- model/model.py:26-27,132-480
    vtc_b200.evaluation.retrieval_evaluation.compute_recall
                                                          <- evaluation/retrieval_evaluation.py:23-47
Everything runs on hand-written sm_100a CUDA kernels behind the C ABI of include/vtc_b200.h;
there is no CPU or eager fallback.
"""
from . import _ffi  # noqa: F401
from ._ffi import VtcError  # noqa: F401

__version__ = "0.1.0"
