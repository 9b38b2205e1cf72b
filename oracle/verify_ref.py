"""TEST INFRASTRUCTURE ONLY (see oracle/README.md): a text-level Python restatement of the read-level
verification that follows the position map in ped.pl's `method=kmer` with a reference,
and of the VCF writer behind it.  It mirrors the Perl line by line — the same split / substr / sort /
join semantics on the same strings — and shares no code with the CUDA path.  Pinned to ped.pl's own
`T.kmer.snp` / `T.kmer.vcf` by tests/test_oracle.py (tests/golden/ref30k_verify.json, made by
scripts/make_golden_verify.py).  Pure-Python loops: small cases only.

  snpMkT         ped.pl:2087-2229     candidates, wild / mutant windows of the target's read length
  sortSeq        ped.pl:2622-2641     (order only; the counts below do not depend on it)
  joinTarget     ped.pl:2601-2620     windows that are reads of the target (`join` with sort_uniq)
  snpMkC         ped.pl:2231-2305     the same with the control's read length
  joinControl    ped.pl:2576-2599
  kmerReadCount  ped.pl:2388-2500     cw cm tw tm per candidate, genotype M / H / R / N, join with the map
  toVcf          ped.pl:1532-1593
"""
from __future__ import annotations

import math
import re

NUC = "ACGT"
TAGS = [a + b + c for a in NUC for b in NUC for c in NUC]
_NUMERIC = re.compile(r"^[0-9]+\n?\Z")


def _f(row, i):
    """Perl: an array element past the end is undef, printed as the empty string."""
    return row[i] if 0 <= i < len(row) else ""


def _num(s: str) -> int:
    """Perl numeric value of a string of the shapes that occur here (digits, possibly none)."""
    m = re.match(r"\s*([+-]?[0-9]+)", s)
    return int(m.group(1)) if m else 0


def _pad_chr(c: str) -> str:
    if _NUMERIC.match(c):
        c = ("000" + c)[-3:]
    return c


def _pad_pos(p: str) -> str:
    return ("000000000000" + p)[-11:]


def _strip_zeros(c: str) -> str:
    return re.sub(r"^0+", "", c)


def snp_list(map_lines):
    """snpMkT up to `$tmpdir/snp.list` (ped.pl:2094-2152): the `chr pos ref alt` strings, sorted, unique."""
    tmp = set()
    for line in map_lines:
        row = line.split()
        row = list(row) + [""] * max(0, 6 - len(row))
        row[0] = _pad_chr(row[0])
        row[1] = _pad_pos(row[1])
        ref, alt = row[4], row[5]
        alt = alt.replace(ref, "", 1) if ref else alt
        tmp.add(f"{row[0]} {row[1]} {ref} {alt}")
    out = []
    for line in sorted(tmp, key=lambda s: s.encode()):      # sort | uniq, bytewise (C locale)
        row = line.split()
        out.append(f"{_f(row, 0)} {_f(row, 1)} {_f(row, 2)} {_f(row, 3)}")
    # `uniq -c` of lines that are unique already unless the re-split merged two of them
    uniq = []
    for s in out:
        if not uniq or uniq[-1] != s:
            uniq.append(s)
    return uniq


def windows(snps, chromosomes, length: int, wild: str, mutant: str):
    """The second half of snpMkT / snpMkC (ped.pl:2153-2200, 2246-2295): for every candidate the `length`
    windows of the reference around it and of the reference with the alternative allele (and up to ten
    neighbouring candidates) put in.  Yields (sequence, "chr pos ref alt kind") for the lines that reach a
    tag file (a window that does not start with three of ACGT goes to a handle that is not open)."""
    prev_chr = None
    dat = []
    seq = None
    for line in snps:
        row = ("1 " + line).split()                       # `uniq -c` puts the count in front
        chr_, pos, ref, alt = _f(row, 1), _f(row, 2), _f(row, 3), _f(row, 4)
        chr_ = _strip_zeros(chr_)
        if prev_chr is None or prev_chr != chr_:
            seq = chromosomes.get("chr" + chr_)           # open(IN, "$refdir/chr$chr"); a miss reads nothing
            dat = []
        prev_chr = chr_
        pos = _num(pos)
        dat.append(f"{pos} {ref} {alt}")                  # $count = int(1 + 25/4) >= 5, always
        if len(dat) - 1 >= 10:
            dat.pop(0)
        if pos - length < 0:
            continue
        ref_seq = seq[pos - length: pos - length + 2 * length - 1] if seq is not None else ""
        if len(ref_seq) != 2 * length - 1:
            continue
        head = ref_seq[:length - 1]
        tail = ref_seq[length: length + length]
        mut_seq = head + alt + tail
        for d in dat:
            r = d.split()
            tpos, ialt = _num(_f(r, 0)), _f(r, 2)
            i = tpos - (pos - length) - 1
            if 0 < i < 2 * length and i != length - 1:
                mut_seq = mut_seq[:i] + ialt + mut_seq[i + 1:]
        for i in range(length):
            tw = ref_seq[i:i + length]
            tm = mut_seq[i:i + length]
            if tw[:3] in TAGS:
                yield tw, f"{chr_} {pos} {ref} {alt} {wild}"
            if tm[:3] in TAGS:
                yield tm, f"{chr_} {pos} {ref} {alt} {mutant}"


def selected(lines, reads):
    """sortSeq + joinTarget / joinControl: `join <(zcat sort_uniq.TAG) <(zcat TAG) | cut -d ' ' -f 2-`.
    A window line joins once with the read equal to its first field; join re-splits the rest on blanks."""
    for seqs, rest in lines:
        if seqs in reads:
            yield " ".join(rest.split())


def _merge_join(a, b):
    """GNU join on the first blank-separated field of two line lists, as a plain merge (which is what it
    does whether or not the inputs are in order); output `key rest-of-1 rest-of-2`, single spaces."""
    ka = [x.split() for x in a]
    kb = [x.split() for x in b]
    i = j = 0
    out = []
    while i < len(ka) and j < len(kb):
        x, y = _f(ka[i], 0).encode(), _f(kb[j], 0).encode()
        if x < y:
            i += 1
        elif x > y:
            j += 1
        else:
            i2 = i
            while i2 < len(ka) and _f(ka[i2], 0).encode() == y:
                i2 += 1
            j2 = j
            while j2 < len(kb) and _f(kb[j2], 0).encode() == x:
                j2 += 1
            for p in range(i, i2):
                for q in range(j, j2):
                    out.append(" ".join([_f(ka[p], 0)] + ka[p][1:] + kb[q][1:]))
            i, j = i2, j2
    return out


def _genotype(cw, cm, tw, tm):
    if cw >= 5 and cm <= 1:
        if tm >= 5 and tw <= 1:
            return "M"
        if (tm + tw) > 0 and 0.3 < tm / (tm + tw) < 0.7 and cm <= 1 and tw >= 5:
            return "H"
        return "R"
    return "N"


def read_count(map_lines, target_hits, control_hits, length: int, clength: int) -> str:
    """kmerReadCount (ped.pl:2388-2500) -> the text of `$target.kmer.snp`."""
    mp = []
    for line in map_lines:
        row = line.split()
        row = list(row) + [""] * max(0, 2 - len(row))
        row[0] = _pad_chr(row[0])
        if _num(_f(row, 1)) < length or _num(_f(row, 1)) < clength:
            continue
        p = _pad_pos(_f(row, 1))
        mp.append(f"{_f(row, 0)}:{p}\t" + "\t".join(_f(row, i) for i in range(2, 15)) + "\n")
    mp.sort(key=lambda s: s.encode())
    count = {}
    for line in list(target_hits) + list(control_hits):
        row = line.split()
        row = row + [""] * max(0, 2 - len(row))
        row[0] = _pad_chr(row[0])
        row[1] = _pad_pos(row[1])
        d = " ".join(row)
        count[d] = count.get(d, 0) + 1
    verify = []
    cm = cw = tm = tw = 0
    prev = ""

    def flush():
        g = _genotype(cw, cm, tw, tm)
        pr = prev.split()
        pr = pr + [""] * max(0, 2 - len(pr))
        pr[0] = _pad_chr(pr[0])
        pr[1] = _pad_pos(pr[1])
        return f"{pr[0]}:{pr[1]}\t{cw}\t{cm}\t{tw}\t{tm}\t{g}\n"

    for key in sorted(count, key=lambda s: s.encode()):
        c = count[key]
        row = key.split()
        row[0] = _strip_zeros(row[0])
        row[1] = str(_num(row[1]))
        d = "\t".join(row[:-1])
        if d != prev:
            line = flush()
            if prev != "":
                verify.append(line)
            cm = cw = tm = tw = 0
        kind = row[-1]
        if kind == "cm":
            cm = c
        if kind == "cw":
            cw = c
        if kind == "tm":
            tm = c
        if kind == "tw":
            tw = c
        prev = d
    verify.append(flush())
    out = []
    for line in _merge_join(mp, verify):
        row = line.split()
        cp = _f(row, 0).split(":")
        chr_ = _strip_zeros(_f(cp, 0))
        pos = _num(_f(cp, 1))
        out.append(f"{chr_}\t{pos}\t" + "\t".join(_f(row, i) for i in range(1, 19)) + "\n")
    return "".join(out)


def first_line_length(sort_uniq: dict) -> int:
    """getLength (ped.pl:3486-3502): the length of the first line of `zcat sort_uniq/*.gz`."""
    for t in TAGS:
        if sort_uniq.get(t):
            return len(sort_uniq[t][0])
    return 0


def verify(map_lines, target_sort_uniq: dict, control_sort_uniq: dict, chromosomes: dict) -> str:
    """map lines (`cat $tmpdir/$target.map.*`), the two sort_uniq line sets (dict tag -> sorted lines) and the
    chromosome files (dict "chrNAME" -> text) -> `$target.kmer.snp` of the run with a reference."""
    length = first_line_length(target_sort_uniq)
    clength = first_line_length(control_sort_uniq)
    t_reads = set(r for v in target_sort_uniq.values() for r in v)
    c_reads = set(r for v in control_sort_uniq.values() for r in v)
    snps = snp_list(map_lines)
    t_hits = list(selected(windows(snps, chromosomes, length, "tw", "tm"), t_reads))
    c_hits = list(selected(windows(snps, chromosomes, clength, "cw", "cm"), c_reads))
    return read_count(map_lines, t_hits, c_hits, length, clength)


def _perl_num(x: float) -> str:
    """Perl prints a number with %.15g."""
    if x == int(x) and abs(x) < 1e15:
        return str(int(x))
    return "%.15g" % x


def to_vcf(kmer_snp: str, target: str) -> str:
    """toVcf, method kmer (ped.pl:1532-1574)."""
    out = ["##fileformat=VCFv4.2\n"
           "##INFO=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n"
           "##INFO=<ID=DP,Number=1,Type=Integer,Description=\"Approximate read depth)\">\n"
           "##INFO=<ID=AF,Number=.,Type=Float,Description=\"Allele Frequency\">\n"
           "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n"
           "##FORMAT=<ID=AD,Number=.,Type=Integer,Description=\"Allelic depths for the reference and alternate "
           "alleles in the order listed\">\n"
           "##FORMAT=<ID=DP,Number=1,Type=Integer,Description=\"Read depth\">\n"
           f"#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t{target}\n"]
    output = ""
    prev = ""
    for line in kmer_snp.split("\n")[:-1] if kmer_snp.endswith("\n") else kmer_snp.split("\n"):
        row = line.split("\t")
        # Perl's split drops trailing empty fields
        while row and row[-1] == "":
            row.pop()
        if len(_f(row, 0)) <= 3 or _NUMERIC.match(_f(row, 0)):
            if row:
                row[0] = "chr" + row[0]
            else:
                row = ["chr"]
        r17, r18 = _num(_f(row, 17)), _num(_f(row, 18))
        dp = r17 + r18
        if dp == 0:
            continue
        af = int(1000 * r18 / dp) / 1000
        try:
            p = (0.001 * dp) ** r18
        except OverflowError:
            p = float("inf")
        if p == 0:
            qual = 1000
        else:
            qual = int((-10 * math.log(p) / math.log(10)) * 1000) / 1000
        if "M" in line:
            output = (f"{row[0]}\t{_f(row, 1)}\t.\t{_f(row, 4)}\t{_f(row, 5)}\t{_perl_num(qual)}\t.\t"
                      f"GT=1/1;AF={_perl_num(af)};DP={_f(row, 18)}\tGT:AD:DP\t1/1:{_f(row, 17)},{_f(row, 18)}:{dp}\n")
        elif "H" in line:
            output = (f"{row[0]}\t{_f(row, 1)}\t.\t{_f(row, 4)}\t{_f(row, 5)}\t{_perl_num(qual)}\t.\t"
                      f"GT=0/1;AF={_perl_num(af)};DP={_f(row, 18)}\tGT:AD:DP\t0/1:{_f(row, 17)},{_f(row, 18)}:{dp}\n")
        if output != prev:
            out.append(output)
        prev = output
    return "".join(out)
