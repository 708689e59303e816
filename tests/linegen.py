"""Seeded generators of pileup text for the parity tests (shared by make_golden.py, CPU and GPU tests).

Three flavours:
  realistic_line   what samtools mpileup emits (SURVEY.md section 8d's mix of markers)
  nasty_line       legal-but-odd bases strings: caret chains, indel tokens of every shape, IUPAC, digits,
                   '>' '<' '#', short / long quality strings, lowercase reference, odd integers
  broken_line      lines on which the reference raises (too few columns, bad integers, blank)
"""
from __future__ import annotations

import random

CHROM = "gi|9626243|ref|NC_001416.1|"
_ACGT = "ACGT"


def _read_char(rng, ref, sub=0.007, n_rate=0.008):
    fwd = rng.random() < 0.5
    x = rng.random()
    if x < sub:
        c = rng.choice([b for b in _ACGT if b != ref.upper()])
    elif x < sub + n_rate:
        c = "N"
    else:
        return "." if fwd else ","
    return c if fwd else c.lower()


def realistic_line(rng, pos, chrom=CHROM, depth=None, ref=None, indel_rate=0.001, del_rate=0.0015,
                   zero_rate=0.0008, alt=None, alt_frac=0.97):
    ref = ref or rng.choice(_ACGT)
    if depth is None:
        depth = min(60, max(0, int(rng.gauss(24, 5))))
    if rng.random() < zero_rate or depth == 0:
        return "%s\t%d\t%s\t0\t*\t*\n" % (chrom, pos, ref)
    bases, quals = [], []
    for _ in range(depth):
        if rng.random() < 1 / 130:
            bases.append("^" + rng.choice("KIUS!~]"))
        if alt is not None and rng.random() < alt_frac:
            bases.append(alt if rng.random() < 0.5 else alt.lower())
        elif rng.random() < del_rate:
            bases.append("*")
        else:
            bases.append(_read_char(rng, ref))
        if rng.random() < indel_rate:
            n = rng.randint(1, 13)
            s = "".join(rng.choice(_ACGT) for _ in range(n))
            bases.append(rng.choice("+-") + str(n) + (s if rng.random() < 0.5 else s.lower()))
        if rng.random() < 1 / 130:
            bases.append("$")
        quals.append(chr(33 + rng.randint(13, 39)))
    return "%s\t%d\t%s\t%d\t%s\t%s\n" % (chrom, pos, ref, depth, "".join(bases), "".join(quals))


_ODD = "ACGTNacgtn.,*" * 3 + "RYKMSWrykmsw><#=0123456789+-^$^$[]_`{}~!\"%&'()/:;?@\\|"


def nasty_line(rng, pos, chrom=CHROM):
    ref = rng.choice("ACGTNacgtnRy*.,")
    depth = rng.choice([1, 2, 3, 5, 8, 13, 30, 70, 150, 300])
    toks = []
    for _ in range(depth):
        x = rng.random()
        if x < 0.55:
            toks.append(rng.choice(".,.,.,ACGTacgtNn*"))
        elif x < 0.65:
            toks.append("^" + rng.choice("^+-$0123456789.,AaKIUS~!*"))
        elif x < 0.72:
            toks.append("$")
        elif x < 0.82:
            n = rng.choice([0, 1, 2, 3, 9, 10, 11, 25, 100, 5000])
            k = n if rng.random() < 0.7 else rng.randint(0, n + 3)
            toks.append(rng.choice("+-") + str(n) + "".join(rng.choice("ACGTNacgtn*.,^$+-12") for _ in range(min(k, 40))))
        elif x < 0.86:
            toks.append(rng.choice("+-"))
        else:
            toks.append(rng.choice(_ODD))
    bases = "".join(toks)
    if rng.random() < 0.15:
        bases += "^"
    nq = len(bases) if rng.random() < 0.3 else rng.randint(1, max(1, 2 * depth))
    quals = "".join(chr(rng.randint(33, 126)) for _ in range(nq))
    raw = rng.choice([str(depth), str(depth), "+%d" % depth, "0%d" % depth, "1_0", "-3"])
    p = rng.choice([str(pos), str(pos), "+%d" % pos, "000%d" % pos, "%d_0" % pos])
    sep = rng.choice(["\t", "\t", "\t", " ", "  \t", "\t \x1c"])
    tail = rng.choice(["\n", "\n", "\r\n", " \n", "\textra\tcols\n", "\x0b\n"])
    lead = rng.choice(["", "", "", " ", "\t"])
    return lead + sep.join([chrom, p, ref, raw, bases, quals]) + tail


def broken_line(rng, pos, chrom=CHROM):
    kind = rng.randrange(9)
    if kind == 0:
        return "\n"
    if kind == 1:
        return "%s\n" % chrom
    if kind == 2:
        return "%s\t%d\n" % (chrom, pos)
    if kind == 3:
        return "%s\t%d\tA\n" % (chrom, pos)
    if kind == 4:
        return "%s\t%dx\tA\t3\t...\tIII\n" % (chrom, pos)
    if kind == 5:
        return "%s\t%d\tA\t3x\t...\tIII\n" % (chrom, pos)
    if kind == 6:
        return "%s\t%d\tA\t3\t...\n" % (chrom, pos)
    if kind == 7:
        return "%s\t_%d\tA\t3\t...\tIII\n" % (chrom, pos)
    return "%s\t%d\tA\t\t\n" % (chrom, pos)


def pileup_text(seed, n_lines, chrom=CHROM, nasty=0.0, start=1, gaps=0.0, sites=None, depth=None):
    """A whole synthetic pileup file.  sites: {pos: alt base} forces a variant pile at those positions."""
    rng = random.Random(seed)
    out = []
    pos = start
    for _ in range(n_lines):
        if rng.random() < nasty:
            out.append(nasty_line(rng, pos, chrom))
        else:
            alt = sites.get(pos) if sites else None
            out.append(realistic_line(rng, pos, chrom, alt=alt, depth=depth))
        pos += 1
        if gaps and rng.random() < gaps:
            pos += rng.randint(1, 50)
    return "".join(out)


def strip_bases(bases):
    """pileup.py:276-325 for well-formed input: '^x' pairs, then indel tokens with their sequences, then '$'."""
    import re
    s = re.sub(r"\^.", "", bases)
    out, i = [], 0
    while i < len(s):
        m = re.match(r"[+-](\d+)", s[i:])
        if m:
            i += len(m.group(0)) + int(m.group(1))
        else:
            out.append(s[i])
            i += 1
    return "".join(out).replace("$", "")


INDEL_PIECES = ["+1A", "-1c", "+2AC", "-3acg", "+10ACGTACGTAC", "-999" + "a" * 999, "+1000" + "C" * 1000, "+", "-", "+0", "-0A",
                "+1", "+2A", "+3AC", "+1.", "-2,.", "+1*", "-2*a", "^+", "^-", "^+1A", "+1^", "-1$", "+1A5", "+01A", "+1A+1C",
                "-1a-2cc", "+1A$", "$+1A", "^K+1A", "+2^K", "*", "+1N", "-1n", "+12ACGTNacgtn*"]


def indel_corner_lines(rng, n):
    """n lines (positions 1..n) of '.'/',' runs with well- and ill-formed indel tokens between, before and behind them,
    so that the tokens meet the column's end at every byte alignment; some with a quality string of another length."""
    lines = []
    for k in range(n):
        ref = rng.choice("ACGT")
        parts = [rng.choice(".,") * rng.randint(0, 9) for _ in range(rng.randint(1, 4))]
        for x in range(len(parts) - (0 if rng.random() < 0.5 else 1)):          # a token between runs, sometimes last
            parts[x] += rng.choice(INDEL_PIECES)
        if rng.random() < 0.3:
            parts.insert(0, rng.choice(INDEL_PIECES))                          # ... and sometimes first
        bases = "".join(parts)
        nb = len(strip_bases(bases)) if rng.random() < 0.8 else rng.randint(0, 12)
        qual = "".join(chr(33 + rng.randint(13, 39)) for _ in range(nb))
        lines.append("%s\t%d\t%s\t%d\t%s\t%s\n" % (CHROM, 1 + k, ref, max(nb, 1), bases, qual))
    return lines
