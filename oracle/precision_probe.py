"""TEST INFRASTRUCTURE ONLY — CPU probe of how much operand precision the KEEP forward needs (no GPU).

Runs the oracle restatement with the conv / linear weights of chosen module groups rounded to fp16 (= what a tensor-core
mode with single-fp16 weights, i.e. fewer than three MMA passes per MAC, computes at best) and reports the decoded-pixel
error against the fp32 run, with the discrete code indices and the flows teacher-forced so a flipped index cannot mask
the continuous error.  Result on the seeded synthetic weights (T = 3): fp16 weights anywhere -- even in the generator
alone -- already cost 1.5e-2 .. 7.7e-2 max-abs on clamped pixels against a bar of 1e-2, which is why the engine's default
tensor-core mode carries weights AND activations as (hi, lo) fp16 pairs.

usage: python oracle/precision_probe.py [T]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import keep_oracle, weights  # noqa: E402

GROUPS = (
    ("every conv / linear weight", lambda k: True),
    ("generator + CFT + CFA", lambda k: k.startswith(("generator.", "cft.", "cfa."))),
    ("LQ encoder + hq_encoder", lambda k: k.startswith(("encoder.", "hq_encoder."))),
    ("code transformer + head", lambda k: k.startswith(("ft_layers.", "feat_emb.", "idx_pred_layer."))),
    ("GMFlow", lambda k: k.startswith("flownet.")),
)


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    torch.set_num_threads(os.cpu_count() or 1)
    sd = weights.make_state_dict(seed=0)
    x = weights.make_clip(T, seed=1234, coherent=True)
    out, cap = keep_oracle.keep_forward(sd, x, capture=True)
    for name, pred in GROUPS:
        sdr = {k: (v.half().float() if v.dim() >= 2 and pred(k) and k not in ("position_emb", "quantize.embedding.weight") else v)
               for k, v in sd.items()}
        flows = None if name == "GMFlow" else cap["flows"]
        o, c = keep_oracle.keep_forward(sdr, x, force_codes=cap["codes"], force_flows=flows, capture=True)
        e = [float((o[:, i].clamp(-1, 1) - out[:, i].clamp(-1, 1)).abs().max()) for i in range(T)]
        print("fp16 weights in %-28s max-abs per frame %s   logits %.2e   flows %.2e"
              % (name, ["%.1e" % v for v in e], float((c["logits"] - cap["logits"]).abs().max()),
                 float((c["flows"] - cap["flows"]).abs().max())), flush=True)


if __name__ == "__main__":
    main()
