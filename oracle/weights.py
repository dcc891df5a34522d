"""Seeded synthetic weights / clips now live in the package (`comfyui-keep_b200/synth.py`) because
bench.py's own arm needs them too; re-exported here for the oracle-side scripts."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import keep_b200  # noqa: E402

_s = sys.modules["comfyui_keep_b200.synth"] if "comfyui_keep_b200.synth" in sys.modules else __import__(
    "importlib").import_module("comfyui_keep_b200.synth")
load_shapes, make_state_dict, make_clip = _s.load_shapes, _s.make_state_dict, _s.make_clip
make_latents, make_vq_case = _s.make_latents, _s.make_vq_case
