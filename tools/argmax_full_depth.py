"""north_star's third parity criterion at FULL depth: loss-based option scoring (`llama/model_my_original_mod.py:375-377` +
`engine.py:88-93`) of the product vs the fp32 oracle (TF32 off) on the same random-init LLaMA-7B-shaped weights (all 32 layers),
N items x 5 options x S = 128. Prints one JSON line: argmax agreement, the per-option normalised-loss error, and how the smallest
best-vs-runner-up margin of the oracle compares with that error (an item whose margin is below the error CAN flip legitimately).

    python tools/argmax_full_depth.py [items=256] [layers=32] [items_per_batch=8]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util_parity import full_depth_argmax_report


def main():
    items = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    layers = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    per = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    print(json.dumps(full_depth_argmax_report(items, layers, per)))


if __name__ == "__main__":
    main()
