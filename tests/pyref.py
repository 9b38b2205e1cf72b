"""A deliberately naive, text-level Python restatement of ped.pl's k-mer path
(SURVEY.md Appendix A), used to cross-check oracle/ped_oracle.c on small inputs.
It mirrors the Perl line by line (strings, dicts, the same split/join semantics)
and shares no code with the oracle or the CUDA path."""
from __future__ import annotations

import re

from collections import Counter

NUC = "ACGT"
TAGS = [a + b + c for a in NUC for b in NUC for c in NUC]
_COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}


def complement(s: str) -> str:  # ped.pl:3464-3484
    return "".join(_COMP.get(c, "N") for c in reversed(s))


def sort_uniq(fastq: str, clipping: int = 0):
    """ped.pl:1420-1530 -> (dict tag -> sorted unique lines, clipping log value)."""
    files = {t: [] for t in TAGS}

    def emit(seq):
        t = seq[:3]
        if t in files:
            files[t].append(seq)

    def emit_both(seq):
        emit(seq)
        emit(complement(seq))

    def clipped(seq, c):
        if len(seq) >= c:
            pos = 0
            while True:
                sub = seq[pos:pos + c]
                if len(sub) < c:
                    break
                pos += c
                emit_both(sub)

    lines = fastq.split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    count = 0
    prev = 0
    flag = False
    hist = Counter()
    total = 0
    read_length = 0
    for line in lines:
        if count == 1 and "N" not in line:
            read_length = len(line)
            hist[read_length] += 1
            if read_length != prev and prev != 0 and not clipping:
                flag = True
            prev = read_length
            if not flag:
                if clipping:
                    clipped(line, clipping)
                else:
                    emit_both(line)
            total += 1
        elif count == 4:
            count = 0
        count += 1
    used = read_length
    if flag:
        s = 0
        c = 0
        for ln in sorted(hist, reverse=True):
            s += hist[ln]
            if int(s * 100 / total) >= 90:
                c = ln
                break
        if c > 200:
            c = int(c / int(c / 100))
        elif c < 75:
            c = 75
        used = c
        count = 0
        for line in lines:
            if count == 1 and "N" not in line:
                clipped(line, c)
            elif count == 4:
                count = 0
            count += 1
    return {t: sorted(set(v)) for t, v in files.items()}, used, flag


def count_table(su, k: int = 20):
    """ped.pl:891-932 + 862-889 -> dict tag -> sorted list of (kmer, count)."""
    cnt = Counter()
    for t in TAGS:
        for read in su[t]:
            for i in range(0, len(read) - k + 1):
                seq = read[i:i + k]
                if "N" not in seq:
                    cnt[seq] += 1
    out = {t: [] for t in TAGS}
    for seq in sorted(cnt):
        out[seq[:3]].append((seq, cnt[seq]))
    return out


def lbc_text(rows, k: int = 20) -> str:
    """ped.pl:833-860 for one tag: %count is uninitialised for the first group."""
    out = []
    count = {}
    prev = ""
    for seq, c in rows:
        head, nuc = seq[:k - 1], seq[k - 1]
        if head != prev and prev != "":
            out.append("%s\t%s\t%s\t%s\t%s\n" % (prev, count.get("A", ""), count.get("C", ""), count.get("G", ""),
                                                 count.get("T", "")))
            count = {"A": 0, "C": 0, "G": 0, "T": 0}
        count[nuc] = c
        prev = head
    out.append("%s\t%s\t%s\t%s\t%s\n" % (prev, count.get("A", ""), count.get("C", ""), count.get("G", ""),
                                         count.get("T", "")))
    return "".join(out)


def _join_fields(line: str):
    """GNU join's blank-separated field split (coreutils join.c, xfields)."""
    i, n = 0, len(line)
    while i < n and line[i] in " \t":
        i += 1
    if i == n:
        return []
    f = []
    while True:
        j = i
        while j < n and line[j] not in " \t":
            j += 1
        f.append(line[i:j])
        if j == n:
            return f
        i = j
        while i < n and line[i] in " \t":
            i += 1
        if i == n:
            f.append("")  # trailing blanks leave one empty field
            return f


def _num(x):
    return int(x) if x not in (None, "") else 0


def join_text(lbc_c: str, lbc_t: str) -> str:
    """ped.pl:703-782 for one tag (`join` + two Perl passes)."""
    a = [l for l in lbc_c.split("\n") if l != "" or False]
    b = [l for l in lbc_t.split("\n") if l != "" or False]
    fa = {}
    for l in a:
        f = _join_fields(l)
        fa[f[0] if f else ""] = f[1:]
    out = []
    for l in b:
        f = _join_fields(l)
        key = f[0] if f else ""
        if key not in fa:
            continue
        joined = " ".join([key] + fa[key] + f[1:])
        row = joined.split()
        row += [None] * (9 - len(row))
        rep = nohita = nohitb = pola = polb = 0
        aa = bb = ""
        for i in range(1, 5):
            cv, tv = _num(row[i]), _num(row[i + 4])
            if cv > 100:
                rep += 1
            elif cv >= 3:
                aa += NUC[i - 1]
                if tv <= 1:
                    pola += 1
            elif cv <= 1:
                nohita += 1
            if tv > 100:
                rep += 1
            elif tv >= 3:
                bb += NUC[i - 1]
                if cv <= 1:
                    polb += 1
            elif tv <= 1:
                nohitb += 1
        if aa != bb and aa and bb and rep == 0 and nohita >= 2 and nohitb >= 2 and (pola == 1 or polb == 1):
            tmp = "\t".join([row[0], aa, bb] + [("" if x is None else x) for x in row[1:9]])
            f2 = tmp.split()
            f2 += [None] * (11 - len(f2))
            name = f2[0]
            c = f2[3:7]
            t = f2[7:11]
            ct = sum(_num(x) for x in c)
            tt = sum(_num(x) for x in t)
            if ct == 0 or tt == 0:
                continue
            cont = "".join(NUC[i] for i in range(4) if _num(c[i]) / ct > 0.25)
            alt = "".join(NUC[i] for i in range(4) if _num(t[i]) / tt > 0.25)
            if cont != alt:
                out.append("\t".join([name, cont, alt] + [("" if x is None else x) for x in c + t]) + "\n")
    return "".join(out)


def run(control_fastq: str, target_fastq: str, k: int = 20, clipping: int = 0):
    res = {}
    for name, fq in (("target", target_fastq), ("control", control_fastq)):  # ped.pl:291-297 order
        su, used, auto = sort_uniq(fq, clipping)
        if auto:
            clipping = used  # `$clipping` is a global in ped.pl
        ct = count_table(su, k)
        res[name] = dict(sort_uniq=su, clipping=used, count=ct, lbc={t: lbc_text(ct[t], k) for t in TAGS})
    # join needs the inputs sorted; the rows already are.  Unpaired keys are dropped.
    res["kmer_snp"] = "".join(join_text(res["control"]["lbc"][t], res["target"]["lbc"][t]) for t in TAGS)
    return res


def cnv_text(ref20_uniq_text: bytes, control_table, target_table, k: int = 20) -> bytes:
    """cnvJoin + cnvJoinSub (ped.pl:934-989), line by line: the full outer join of the two count tables
    (a missing count is 0), inner-joined with the unique reference k-mers, `chr pos strand t c`, then
    `sort -k 1 -n -k 2 -n` under LANG=C (numeric prefix of the chromosome name, position, whole line).
    control_table / target_table: (kmers u64 array, counts u32 array)."""
    code = {"A": 0, "C": 1, "G": 2, "T": 3}

    def enc(s):
        v = 0
        for ch in s:
            v = (v << 2) | code[ch]
        return v

    ct = dict(zip((int(x) for x in control_table[0]), (int(x) for x in control_table[1])))
    tt = dict(zip((int(x) for x in target_table[0]), (int(x) for x in target_table[1])))

    def num(s):
        m = re.match(r"[ \t]*(-?[0-9]*\.?[0-9]*)", s)
        try:
            return float(m.group(1))
        except ValueError:
            return 0.0

    out = []
    for line in ref20_uniq_text.decode().splitlines():
        km, chr_, pos, strand = line.split()
        key = enc(km)
        if key not in ct and key not in tt:
            continue
        out.append(f"{chr_}\t{pos}\t{strand}\t{tt.get(key, 0)}\t{ct.get(key, 0)}\n")
    out.sort(key=lambda ln: (num(ln.split("\t")[0]), num(ln.split("\t")[1]), ln.encode()))
    return "".join(out).encode()
