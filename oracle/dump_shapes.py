"""TEST INFRASTRUCTURE ONLY — dump the reference's own `KEEP(**cfg).state_dict()` key -> shape tables.

Run in the build container (needs /root/reference):   python oracle/dump_shapes.py

Writes comfyui-keep_b200/keep_state_shapes.json ('KEEP' general config, modules/utils.py:42-57) and
comfyui-keep_b200/keep_state_shapes_asian.json ('Asian' config, modules/utils.py:58-73), in the reference's
state_dict order.  The host mirror's strict `load_state_dict` and the seeded synthetic weights are built on them.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402

OUT = {"KEEP": "keep_state_shapes.json", "Asian": "keep_state_shapes_asian.json"}


def main():
    for name, fn in OUT.items():
        net = ref_loader.load_reference_keep(config=name)
        shapes = {k: list(v.shape) for k, v in net.state_dict().items()}
        path = os.path.join(ROOT, "comfyui-keep_b200", fn)
        if os.path.exists(path):
            with open(path) as f:
                old = json.load(f)
            print(name, "matches the committed table:", old == shapes and list(old) == list(shapes))
        with open(path, "w") as f:
            json.dump(shapes, f, indent=0)
        print(name, len(shapes), "tensors,", sum(int(__import__("math").prod(s)) for s in shapes.values()) / 1e6, "M parameters")


if __name__ == "__main__":
    main()
