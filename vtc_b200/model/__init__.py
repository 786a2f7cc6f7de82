from .lazy import LazySim  # noqa: F401
from .loss import clip_loss  # noqa: F401
from .metric import (BaseMetric, LossMetric, MetricTracker, RecallAtK,  # noqa: F401
                     ScalarPerBatchMetric)
from .model import (CAMTransformer, PretrainedCLIP, PretrainedCLIP_finaltf,  # noqa: F401
                    PretrainedCLIP_TimeSformer, PretrainedCLIP_TimeSformer_finaltf,
                    PretrainedCLIPBase, normalize)
