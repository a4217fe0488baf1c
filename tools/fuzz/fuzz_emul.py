"""Fuzz the kernels on the CPU: random picture geometries / bit depths / chroma formats / tool mixes, every picture through
oracle/_ref/libb200hevc_emul.so (kernels.cu + engine.cu as warp-lockstep fibers, tests/emul/warp/) and through the oracle;
any differing sample is printed with the parameters that reproduce it.  TEST INFRASTRUCTURE; needs no GPU.

    python tools/fuzz/fuzz_emul.py [--seconds 300] [--seed 1] [--asan]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=300)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--asan", action="store_true", help="use the address sanitizer build (run with LD_PRELOAD=libasan.so)")
    a = ap.parse_args()
    from openhevc_b200 import _lib
    _lib.LIB_PATH = os.path.join(ROOT, "oracle", "_ref", "libb200hevc_emul_asan.so" if a.asan else "libb200hevc_emul.so")
    import numpy as np
    import oracle_lib
    from openhevc_b200 import FrameEngine
    from openhevc_b200.synth import FrameSynth, smooth_frame

    def run_sequence(w, h, cfi, bd, seeds, log2_ctb, **kw):
        """I picture -> slot 0, then pictures predicted from everything decoded so far; every picture compared (tests/test_parity_gpu.py)"""
        eng = FrameEngine(w, h, cfi, bd, log2_ctb_size=log2_ctb, n_slots=4)
        dpb = [[np.zeros_like(p) for p in smooth_frame(w, h, cfi, bd, 0)] for _ in range(4)]
        try:
            for k, seed in enumerate(seeds):
                blob, _ = FrameSynth(w, h, cfi, bd, log2_ctb=log2_ctb, seed=seed, refs=list(range(k)), cur_slot=k, poc=k, **kw).generate()
                oracle_lib.check_decode_order(blob)
                got = eng.decode(blob)
                want = oracle_lib.execute(blob, dpb)
                for p in range(3):
                    diff = np.argwhere(got[p] != want[p])
                    assert len(diff) == 0, f"picture {k} plane {p}: {len(diff)} samples differ, first at (y,x)={tuple(diff[0])}"
                dpb[k] = [p.copy() for p in want]
        finally:
            eng.close()
    rng = np.random.default_rng(a.seed)
    t0, n, bad = time.time(), 0, 0
    while time.time() - t0 < a.seconds:
        cfi = int(rng.choice([1, 1, 2, 3]))
        bd = int(rng.choice([8, 10, 10, 12, 9]))
        log2_ctb = int(rng.choice([4, 5, 6, 6]))
        w = 8 * int(rng.integers(2, 40))
        h = 8 * int(rng.integers(2, 24))
        kw = dict(exotic=float(rng.choice([0, 0.04, 0.3])), weighted=bool(rng.integers(2)), sao_restore=bool(rng.integers(2)),
                  cip=bool(rng.integers(3) == 0), p_intra=float(rng.choice([0.02, 0.12, 0.5, 1.0])), split_bias=float(rng.choice([0.4, 1.0, 2.0])),
                  max_mv=int(rng.choice([8, 64, 300])), bi_frac=float(rng.choice([0.0, 0.6, 1.0])), coded_frac=float(rng.choice([0.1, 0.7, 1.0])),
                  qp=int(rng.integers(10, 50)))
        seeds = [int(s) for s in rng.integers(1, 1 << 30, size=3)]
        desc = f"w={w} h={h} cfi={cfi} bd={bd} log2_ctb={log2_ctb} seeds={seeds} {kw}"
        try:
            run_sequence(w, h, cfi, bd, seeds, log2_ctb, **kw)
        except Exception as e:                                       # noqa: BLE001 -- report and go on
            bad += 1
            print("FAIL", desc, "->", str(e)[:300], flush=True)
        n += 1
    print(f"{n} sequences, {bad} failures, {time.time() - t0:.0f} s")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
