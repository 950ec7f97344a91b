#!/usr/bin/env python3
"""Small, quiet driver for ncu: builds the headline model (K=3965, T=256) and runs a few decodes of
one plan.  No oracle, no CPU baseline — keep the profiled process short."""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "flash-viterbi_b200" / "host"))

import numpy as np  # noqa: E402

from __graft_entry__ import load_pkg  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--engine", default="persistent", choices=["step", "persistent", "sparse"])
    ap.add_argument("--segments", type=int, default=64)
    ap.add_argument("--iters", type=int, default=4)
    ap.add_argument("--K", type=int, default=3965)
    ap.add_argument("--T", type=int, default=256)
    ap.add_argument("--beam", type=int, default=0)
    ap.add_argument("--batch", type=int, default=1)
    a = ap.parse_args()
    import gen_hmm

    fv = load_pkg()
    A, B, Pi = gen_hmm.make_hmm(a.K, 50, 0.112, 1)
    f = gen_hmm.as_reference_floats
    ctx = fv.Context(0)
    model = fv.Model(ctx, f(A), f(B), f(Pi))
    eng = {"step": fv.ENGINE_STEP, "persistent": fv.ENGINE_PERSISTENT, "sparse": fv.ENGINE_SPARSE}[a.engine]
    plan = fv.Plan(model, a.T, a.segments, a.batch, a.beam, eng)
    obs = np.stack([gen_hmm.observations(a.T, 50, 1000 + b) for b in range(a.batch)])
    plan.upload(obs)
    for _ in range(a.iters):
        plan.run()
    paths, scores = plan.download()
    rep = plan.report()
    print(f"engine={a.engine} N={a.segments} decode_ms={rep.decode_ms:.3f} first_pass_ms={rep.first_pass_ms:.3f} "
          f"launches={rep.kernel_launches} score={scores[0]}")


if __name__ == "__main__":
    main()
