"""Timeline of the BPTT kernels (end times per kernel and step on one clock), IC3_BPTT_TRACE=1.
IC3_BPTT_TRACE=1 python profiles/tools/bptt_trace.py"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "profiles", "tools"))
os.environ["IC3_BPTT_TRACE"] = "1"
import bptt_profile  # noqa: E402


def main():
    sys.argv = [sys.argv[0], "pp_hard_ic3net", "24"]
    bptt_profile.main()
    lib = C.CDLL(os.path.join(ROOT, "ic3net_b200", "libic3net_b200.so"))
    out = (C.c_float * (16 * 7))()
    print("rc", lib.ic3_debug_bptt_trace(out))
    names = ["heads", "prep", "scale", "gates", "dgrad", "comm", "wgrad"]
    print("end times in us after ic3_bptt_begin; step index = T-1-t")
    print("step " + " ".join("%9s" % n for n in names))
    for i in range(16):
        print("%4d " % i + " ".join("%9.1f" % out[i * 7 + k] for k in range(7)))


if __name__ == "__main__":
    main()
