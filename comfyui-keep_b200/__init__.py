"""keep_b200 — Blackwell-native (sm_100a) inference path for KEEP, drop-in at `keep_net(clip)`.

The directory name carries a hyphen (repo layout contract), so import it through the top-level
shim:  `import keep_b200`  (see /keep_b200.py) or `importlib` with this file's path.
"""
from .keep_net import (KeepNetB200, KEEP_GENERAL_CFG, KEEP_ASIAN_CFG, install_into_model_pack, install_into_loader, from_reference,  # noqa: F401
                       DEFAULT_FLAGS, lib_path,
                       vector_quantize)
from .build import build  # noqa: F401
from . import synth  # noqa: F401,E402
from . import sharding  # noqa: F401,E402
