"""Drop-in shim: with `attention-based-e2e-asr-dnn_b200/` on sys.path ahead of the reference checkout, the reference's
`from src.models import ListenAttendSpell` / `from src.modules import ...` (src/train.py:22, src/infer.py:14,
src/lmtrain.py) resolve to the B200-native modules.  See INTEGRATION.md."""
