"""Drop-in for the model class of the reference's src/lmtrain.py (Rewriter, :95-253).  The reference's own lmtrain.main
cannot construct its Trainer, so only the model is mirrored here; tests/test_rewriter.py drives a train step."""
from las_b200.lm import Rewriter  # noqa: F401
from las_b200.models import MultiheadCrossAttention  # noqa: F401
from las_b200.modules import AutoRegDecoderLSTMCell, LockedLSTM  # noqa: F401
