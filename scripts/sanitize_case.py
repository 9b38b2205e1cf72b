#!/usr/bin/env python
"""One test case through the C ABI (ingest, count, join, table, f1 index for reference cases), checked against
the oracle -- small enough to run under compute-sanitizer (scripts/gpu_sanitizer.sh).
usage: python scripts/sanitize_case.py case [case ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from ped_b200 import api  # noqa: E402
from tests.cases import CASES, REF_CASES, make_inputs, make_ref_inputs  # noqa: E402


def main():
    ok = True
    with api.Context(0) as ctx:
        for name in sys.argv[1:]:
            is_ref = name in REF_CASES
            if is_ref:
                fc, ft = make_ref_inputs(REF_CASES[name])
                oc, ot, osnp = oracle.run_ref(fc, ft, 20)
                clip = 0
            else:
                case = CASES[name]
                fc, ft = make_inputs(case)
                clip = case.get("clipping", 0)
                oc, ot, osnp = oracle.run(fc, ft, 20, clip)
            c, t = api.Sample(ctx), api.Sample(ctx)
            c.add_fastq(fc)
            t.add_fastq(ft)
            ti = t.ingest(clip)
            if is_ref:
                c.ingest_reference()
                c.ref20_uniq(20)
                for tag in range(64):
                    ok &= c.ref20_uniq_text(tag) == oracle.ref20_uniq_text(fc, tag)
            else:
                c.ingest(ti.clipping_used if ti.auto_clipped else clip)
            c.count(20)
            t.count(20)
            r = ctx.join(c, t)
            good = r.text() == osnp
            for s, o in ((c, oc), (t, ot)):
                km, ct = s.table()
                okm, oct_ = o.table()
                good = good and km.shape == okm.shape and bool((km == okm).all()) and bool((ct == oct_).all())
            print(f"[{name}] edges {r.num_edges}: {'ok' if good else 'MISMATCH'}", flush=True)
            ok &= good
            r.destroy(); c.destroy(); t.destroy()
    print("SANITIZE_CASE", "PASS" if ok else "FAIL")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
