"""Drop-in surface of the reference's ``DeepLabV3Plus-Pytorch/`` sub-project for the DML hot path:
the embedding models' distance head (``network/utils.py``), the criterion (``utils/loss.py``),
``StreamSegMetrics`` (``metrics/stream_metrics.py``) and the NPM / PLM evaluation steps of
``test_embedding.py`` / ``test_self_distillation.py``."""
from . import baseline, evaluation, loss, metrics, network  # noqa: F401
from .loss import CrossEntropyLoss, CrossEntropyLoss_dis, FocalLoss  # noqa: F401
from .metrics import StreamSegMetrics  # noqa: F401
from .network import (_SimpleSegmentationModel_embedding,  # noqa: F401
                      _SimpleSegmentationModel_embedding_self_distillation)
