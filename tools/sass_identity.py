"""Which kernels of libbarcode_b200 changed at the instruction level between two source states?

A refactor that is meant to leave the shipped kernels alone (a new opt-in variant, a template parameter with
a default) can be checked WITHOUT a GPU: compile the .cu files of both states for sm_100a and compare, kernel
by kernel, the SASS instruction streams (opcodes and operands; addresses and encodings dropped).

    python tools/sass_identity.py save /tmp/before      # at the old state (e.g. after `git stash`)
    python tools/sass_identity.py diff /tmp/before      # at the new state: lists identical / changed / new / gone
"""
from __future__ import annotations

import hashlib
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from barcode_b200 import build as b  # noqa: E402


def kernel_hashes():
    out = {}
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    for src in b.CU_SOURCES:
        obj = "/tmp/sass_identity_%d_%s.o" % (os.getpid(), src)
        subprocess.run([nvcc] + b.NVCC_FLAGS + ["-c", os.path.join(b.CSRC, src), "-o", obj], check=True,
                       capture_output=True)
        sass = subprocess.run(["cuobjdump", "-sass", obj], check=True, capture_output=True, text=True).stdout
        os.unlink(obj)
        cur = None
        for line in sass.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                cur = src + ":" + m.group(1)
                out[cur] = hashlib.sha1()
                continue
            m = re.search(r"/\*[0-9a-f]{4}\*/\s+(.*?);", line)
            if cur and m:
                out[cur].update(m.group(1).encode())
    return {k: v.hexdigest() for k, v in out.items()}


def main():
    if len(sys.argv) != 3 or sys.argv[1] not in ("save", "diff"):
        raise SystemExit(__doc__)
    path = sys.argv[2]
    now = kernel_hashes()
    if sys.argv[1] == "save":
        with open(path, "w") as f:
            json.dump(now, f)
        print("saved %d kernels to %s" % (len(now), path))
        return
    with open(path) as f:
        old = json.load(f)
    same = [k for k in old if now.get(k) == old[k]]
    changed = [k for k in old if k in now and now[k] != old[k]]
    gone = [k for k in old if k not in now]
    new = [k for k in now if k not in old]
    print("%d identical, %d changed, %d new, %d gone" % (len(same), len(changed), len(new), len(gone)))
    names = subprocess.run(["c++filt"], input="\n".join(k.split(":", 1)[1] for k in changed + new + gone),
                           capture_output=True, text=True).stdout.splitlines()
    for tag, k, n in zip(["CHANGED"] * len(changed) + ["NEW"] * len(new) + ["GONE"] * len(gone), changed + new + gone, names):
        print("%-8s %s  %s" % (tag, k.split(":", 1)[0], re.sub(r"\(.*$", "", n)[:120]))
    sys.exit(1 if changed or gone else 0)


if __name__ == "__main__":
    main()
