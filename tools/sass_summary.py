"""Per-kernel counts of the Blackwell-native SASS mnemonics in the in-tree libraries (cuobjdump -sass; no GPU needed):
UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA bulk tensor load / store), HMMA (mma.sync).

    python tools/sass_summary.py > profiles/r2_sass_summary.txt
"""
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flipped_vqa_b200 import build

MNEMONICS = ("UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMALDG.*MULTICAST", "UTMASTG", "HMMA", "SYNCS", "UTCBAR")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"\(.*", "", n).replace("fvqa::", "") for n in out]


def main():
    exe = "/usr/local/cuda/bin/cuobjdump"
    for variant in ("fp16", "bf16"):
        lib = build.lib_path(variant)
        sass = subprocess.run([exe, "-sass", lib], capture_output=True, text=True).stdout
        kernels, cur = {}, None
        for line in sass.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                cur = kernels.setdefault(m.group(1), [])
            elif cur is not None:
                cur.append(line)
        names = list(kernels)
        pretty = demangle(names)
        print(f"# {os.path.basename(lib)} ({variant} operands): {len(names)} kernels, sm_100a SASS, counts of instructions per kernel")
        print(f"{'kernel':78s} " + " ".join(f"{m[:14]:>14s}" for m in MNEMONICS))
        for n, p in sorted(zip(names, pretty), key=lambda t: t[1]):
            body = "\n".join(kernels[n])
            counts = [len(re.findall(r"\b" + m.replace(".", r"\.").replace(r"\.*", ".*") + r"\b", body)) for m in MNEMONICS]
            print(f"{p[:78]:78s} " + " ".join(f"{c:14d}" for c in counts))
        print()


if __name__ == "__main__":
    main()
