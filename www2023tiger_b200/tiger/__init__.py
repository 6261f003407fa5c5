"""Drop-in mirror of the reference's `tiger` package surface (tiger/data + tiger/model + eval_utils
+ utils) for the per-batch temporal memory path, backed by libtiger_b200.so.

Put `<repo>/www2023tiger_b200` in front of the reference checkout on sys.path and the reference's
own `init_utils.py`, `train_self_supervised.py` and `train_self_supervised_ddp.py` import these
classes instead of theirs.  Class names, constructor keywords, method
signatures, error behaviour and state_dict keys follow the reference; the bodies are new.
"""
import os as _os
import sys as _sys

_ROOT = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _ROOT not in _sys.path:            # `tiger` may be imported as a top-level package: make the kernels' package reachable
    _sys.path.insert(0, _ROOT)
