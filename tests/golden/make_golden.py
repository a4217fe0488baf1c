#!/usr/bin/env python
"""Generates the committed golden fixtures: small synthetic pictures executed by the UNMODIFIED reference's own C
functions (oracle/_ref/libohevc_ref.so through oracle/replay_ref.c).  Run in the build container (needs /root/reference
to have been compiled by oracle/build_ref.sh):  python tests/golden/make_golden.py
Each fixture = the work-list blob, the DPB contents it reads (generator seeds) and the reconstructed picture."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib  # noqa: E402
from openhevc_b200.synth import FrameSynth, smooth_frame  # noqa: E402

CASES = [  # name, w, h, cfi, bd, refs, kwargs
    ("i_420_8", 128, 64, 1, 8, [], dict(exotic=0.05)),
    ("b_420_10_weighted", 192, 128, 1, 10, [1, 2], dict(weighted=True, max_mv=100)),
    ("b_422_10", 128, 128, 2, 10, [1, 2], dict(sao_restore=True)),
    ("p_444_8", 128, 64, 3, 8, [2], dict(exotic=0.05)),
    ("b_420_12", 128, 64, 1, 12, [1, 2], {}),
]


def main():
    assert oracle_lib.ref_lib() is not None, "reference build missing"
    for name, w, h, cfi, bd, refs, kw in CASES:
        blob, _ = FrameSynth(w, h, cfi, bd, seed=sum(map(ord, name)), refs=refs, cur_slot=0, **kw).generate()
        dpb = [smooth_frame(w, h, cfi, bd, 900 + k) for k in range(3)]
        out = oracle_lib.ref_execute(blob, dpb)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), blob=blob, geom=np.array([w, h, cfi, bd]), dpb_seeds=np.array([900, 901, 902]),
                            y=out[0], cb=out[1], cr=out[2])
        print(name, os.path.getsize(os.path.join(HERE, name + ".npz")))


if __name__ == "__main__":
    main()
