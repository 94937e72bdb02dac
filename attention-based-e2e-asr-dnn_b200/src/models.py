from las_b200.models import Listener, ListenAttendSpell, MultiheadCrossAttention, Speller  # noqa: F401
from las_b200.modules import AutoRegDecoderLSTMCell, LockedLSTM, pyramLockedLSTM  # noqa: F401
