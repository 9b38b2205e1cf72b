"""Launched by torchrun (one rank per GPU): the sharded path must reproduce the oracle.

Rank r ingests a contiguous slice of the records of both samples; reads, canonical k-mers
and table rows change hands inside libpedk (peer stores over NVLink, NCCL for the small
collectives).  The tables come back in rank order = key order, so their concatenation over the
ranks must equal the single-process oracle; every rank holds the edges of the (k-1)-mers whose
hash it owns, sorted, so the merged edge lines must equal the oracle's edge list.
usage: torchrun --nproc-per-node N tests/multi_rank_check.py case [case ...]
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from ped_b200 import api  # noqa: E402
from tests.cases import CASES, REF_CASES, make_inputs, make_ref_inputs  # noqa: E402


def slice_records(fq: bytes, rank: int, world: int) -> bytes:
    lines = fq.split(b"\n")
    tail = lines[-1]
    lines = lines[:-1]
    n_rec = (len(lines) + 3) // 4
    a, b = n_rec * rank // world, n_rec * (rank + 1) // world
    part = lines[4 * a:4 * b]
    out = b"".join(l + b"\n" for l in part)
    if rank == world - 1 and tail:
        out += tail
    return out


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = api.Context(local)
    idt = torch.zeros(api.COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(ctx.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    ctx.comm_init(idt.cpu().numpy().tobytes(), rank, world)
    ok = True
    for name in sys.argv[1:]:
        if name.endswith("_cnv"):
            # row f3 on several ranks: the reference (whole, on every rank) indexed, both samples sliced and
            # counted, every rank looks the index up in its key range of the rows, the counts are summed
            import hashlib
            import json
            from tests.cases import make_cnv_inputs
            fa, fc, ft = make_cnv_inputs(REF_CASES[name[:-4]])
            r, c, t = api.Sample(ctx), api.Sample(ctx), api.Sample(ctx)
            r.add_fastq(fa)
            r.ingest_reference()
            r.ref20_uniq(20)
            for smp, fq in ((t, ft), (c, fc)):
                smp.add_fastq(slice_records(fq, rank, world))
                smp.ingest()
                smp.count(20)
            text = r.cnv_text(c, t)
            if rank == 0:
                g = json.load(open(os.path.join(ROOT, "tests", "golden", name + ".json")))
                good = (text.count(b"\n"), hashlib.sha256(text).hexdigest()) == (g["cnv_rows"], g["cnv_sha256"])
                print(f"[{name}] cnv lines {text.count(b'%c' % 10)} vs ped.pl's {g['cnv_rows']}: {'ok' if good else 'MISMATCH'}")
                ok &= good
            r.destroy(); c.destroy(); t.destroy()
            continue
        is_ref = name in REF_CASES
        if is_ref:
            # the control is the reference tiled into reads: every rank is fed the whole FASTA
            # and keeps the windows it owns; the target's records are sliced as usual
            fc, ft = make_ref_inputs(REF_CASES[name])
            clip = 0
        else:
            case = CASES[name]
            fc, ft = make_inputs(case)
            clip = case.get("clipping", 0)
        c, t = api.Sample(ctx), api.Sample(ctx)
        c.add_fastq(fc if is_ref else slice_records(fc, rank, world))
        t.add_fastq(slice_records(ft, rank, world))
        ti = t.ingest(clip)
        ci = c.ingest_reference() if is_ref else c.ingest(clip)
        c.count(20)
        t.count(20)
        r = ctx.join(c, t)
        ck, cc = c.table()
        tk, tc = t.table()
        parts = [None] * world
        dist.all_gather_object(parts, dict(text=r.text(), ck=ck, cc=cc, tk=tk, tc=tc, cd=c.table_digest(),
                                           td=t.table_digest(), cu=int(ci.unique_reads), tu=int(ti.unique_reads)))
        lines = [ln for p in parts for ln in p["text"].split(b"\n") if ln]
        text = b"".join(ln + b"\n" for ln in sorted(lines))  # a line starts with its (k-1)-mer: sorted = key order
        verified = None
        if is_ref:
            # rows f1 + f2 on several ranks: every rank indexes the (whole) reference, maps the merged edge
            # list, and looks every window up among ITS share of the reads; the counters are summed over ranks
            c.ref20_uniq(20)
            map_text = c.map_kmer_text(text)
            verified = c.verify_text(c, t, map_text)
        if rank == 0:
            oc, ot, osnp = oracle.run_ref(fc, ft, 20) if is_ref else oracle.run(fc, ft, 20, clip)
            for key, cnt, o in (("ck", "cc", oc), ("tk", "tc", ot)):
                km = np.concatenate([p[key] for p in parts])
                ct = np.concatenate([p[cnt] for p in parts])
                okm, oct_ = o.table()
                good = km.shape == okm.shape and (km == okm).all() and (ct == oct_).all()
                dg = [p[key[0] + "d"] for p in parts]  # (rows, digest) of every rank's unsorted rows
                good = good and (sum(d[0] for d in dg), sum(d[1] for d in dg) % 2**64) == api.table_digest(okm, oct_)
                print(f"[{name}] {key[0]} table rows {len(km)} vs oracle {len(okm)}: {'ok' if good else 'MISMATCH'}")
                ok &= bool(good)
            uc = sum(p["cu"] for p in parts), sum(p["tu"] for p in parts)
            good = uc == (oc.num_reads, ot.num_reads)
            print(f"[{name}] unique reads {uc} vs oracle {(oc.num_reads, ot.num_reads)}: {'ok' if good else 'MISMATCH'}")
            ok &= good
            good = text == osnp
            print(f"[{name}] edges {text.count(b'%c' % 10)} vs oracle {osnp.count(b'%c' % 10)}: {'ok' if good else 'MISMATCH'}")
            ok &= good
            if verified is not None:
                import json
                gpath = os.path.join(ROOT, "tests", "golden", name + "_verify.json")
                if os.path.exists(gpath):
                    want = json.load(open(gpath))["control_ref"]["kmer_snp"]
                    good = verified.decode() == want
                    print(f"[{name}] verified candidates {verified.count(b'%c' % 10)} lines vs ped.pl's "
                          f"{want.count(chr(10))}: {'ok' if good else 'MISMATCH'}")
                    ok &= good
        r.destroy(); c.destroy(); t.destroy()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    ctx.close()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_RANK_CHECK", "PASS" if ok else "FAIL")
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
