from las_b200.modules import AutoRegDecoderLSTMCell, LockedLSTM, pyramLockedLSTM  # noqa: F401
