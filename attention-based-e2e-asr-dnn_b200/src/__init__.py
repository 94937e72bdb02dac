"""Drop-in shim: with `attention-based-e2e-asr-dnn_b200/` on sys.path ahead of the reference checkout, the reference's
`from src.models import ListenAttendSpell` / `from src.modules import ...` (src/train.py:22, src/infer.py:14,
src/lmtrain.py) resolve to the B200-native modules, while every other `src.*` module (train, infer, utils, constants ...)
still comes from the reference checkout named by $LAS_REFERENCE_SRC.  See INTEGRATION.md."""
import os as _os

_ref = _os.environ.get('LAS_REFERENCE_SRC')
if _ref and _os.path.isdir(_ref) and _ref not in __path__:
    __path__.append(_ref)          # ours first: src.models / src.modules are the B200 ones
