"""Import shim: the package directory is `comfyui-keep_b200/` (hyphenated by the repo layout contract),
which Python cannot import by name — load it under the module name `comfyui_keep_b200`."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "comfyui-keep_b200")
_NAME = "comfyui_keep_b200"

if _NAME in sys.modules:
    pkg = sys.modules[_NAME]
else:
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_PKG_DIR, "__init__.py"),
                                                   submodule_search_locations=[_PKG_DIR])
    pkg = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = pkg
    _spec.loader.exec_module(pkg)

KeepNetB200 = pkg.KeepNetB200
KEEP_GENERAL_CFG = pkg.KEEP_GENERAL_CFG
KEEP_ASIAN_CFG = pkg.KEEP_ASIAN_CFG
vector_quantize = pkg.vector_quantize
install_into_model_pack = pkg.install_into_model_pack
install_into_loader = pkg.install_into_loader
from_reference = pkg.from_reference
DEFAULT_FLAGS = pkg.DEFAULT_FLAGS
build = pkg.build
lib_path = pkg.lib_path
keep_net = sys.modules[_NAME + ".keep_net"]
synth = pkg.synth
sharding = pkg.sharding
