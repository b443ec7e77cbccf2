"""selfmask_b200 — B200-native SelfMask inference + evaluation hot path (hand-written sm_100a CUDA behind a
C-ABI; see include/selfmask_b200.h).  Importing the package does not need a GPU; calling into
it does, and it fails loudly when libselfmask_b200.so is missing — there is no CPU fallback."""
from ._lib import LIB_PATH, SmkConfig, SmkError, lib  # noqa: F401
from .evaluator import BatchRecords, Evaluator, eval_batch, objectness_top1_ties, summarize  # noqa: F401
from .metrics import (AverageMeter, FMeasure, SMeasure, compute_iou, compute_mae, compute_pixel_accuracy,  # noqa: F401
                      finalize, running_mean)
from .model import SelfMaskB200, get_model, weight_table  # noqa: F401
from .parallel import RecordExchange, allreduce_records, gather_records, shard_range  # noqa: F401
from . import synthetic  # noqa: F401
from .datasets import SaliencyFolder, get_dataset  # noqa: F401
