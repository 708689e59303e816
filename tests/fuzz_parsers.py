"""Manual fuzz loop (not collected by pytest): the kernel's three line parsers compiled for the host (csrc/cpu_sim.cpp)
against the oracle on indel-corner lines, realistic lines with many indel tokens, lines with a mutated bases column
and linegen's nasty lines (IUPAC, '>' / '<', "^x" chains, short quality strings, CRLF, no final newline).
    python tests/fuzz_parsers.py [seconds] [first_seed]
Last runs of round 1: 600 s, 12 847 texts x 400 lines of the first three kinds and 480 s, 11 858 texts x 300 nasty
lines: no mismatch.  Round 2, final parsers (byte >= 0x80 ends a column, ST_SIGN / ST_ZERO outcomes, token rounds in the
second look): 600 s from seed 31337000, 10 937 texts x 400 lines, no mismatch."""
import ctypes
import os
import random
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linegen
import test_cpu_sim as t
from oracle import oracle as orc


def main(seconds, first_seed):
    subprocess.check_call(["make", "-s", "-C", t.CSRC, "cpusim"])
    L = ctypes.CDLL(t.SO)
    vp, sz = ctypes.c_void_p, ctypes.c_size_t
    L.cpusim_pileup.restype = ctypes.c_int
    L.cpusim_pileup.argtypes = [vp, sz, ctypes.c_char_p, vp, ctypes.c_int32, vp, vp, sz, vp, vp, sz,
                                ctypes.POINTER(t.CallParams), ctypes.c_int, ctypes.c_int, vp, vp, sz, vp]
    t0, it, n = time.time(), 0, 400
    while time.time() - t0 < seconds:
        seed = first_seed + it
        it += 1
        rng = random.Random(seed)
        kind = it % 4
        if kind == 0:
            lines = linegen.indel_corner_lines(rng, n)
        elif kind == 1:
            rate = rng.choice([0.02, 0.1, 0.3])
            lines = [linegen.realistic_line(rng, 1 + k, indel_rate=rate) for k in range(n)]
        elif kind == 3:
            text = linegen.pileup_text(seed, n, nasty=rng.choice([0.1, 0.5, 1.0])).encode()
            lines = [ln.decode("latin-1") + "\n" for ln in text.split(b"\n")[:-1]]
        else:
            lines = [linegen.realistic_line(rng, 1 + k, indel_rate=0.05) for k in range(n)]
            for k in range(0, n, 7):                                   # one byte of the bases column replaced
                f = lines[k].split("\t")
                b = list(f[4])
                if b:
                    b[rng.randrange(len(b))] = rng.choice("+-^$*.,ACGTNacgtn0123456789")
                f[4] = "".join(b)
                lines[k] = "\t".join(f)
        ps = t.PARAM_SETS[it % len(t.PARAM_SETS)]
        op = orc.make_params(*ps)
        good = []
        try:
            for k, line in enumerate(lines):                           # lines the reference stops at: one by one
                try:
                    orc.pileup_consensus(line.encode("latin-1"), [], [], op, parse_all=True, want_lines=True)
                    good.append(line)
                except orc.OracleError:
                    t._compare(L, line.encode("latin-1"), [(linegen.CHROM, 1 + k)], [], ps, True)
            text = "".join(good).encode("latin-1")
            if kind == 3 and it % 8 == 3:
                text = text.replace(b"\r\n", b"\n").replace(b"\n", b"\r\n")
            snps = [(linegen.CHROM, p) for p in sorted(rng.sample(range(1, n + 1), 80))]
            for all_pos in (False, True):
                t._compare(L, text, snps, [], ps, all_pos)
        except AssertionError:
            print("MISMATCH: seed %d, kind %d" % (seed, kind))
            raise
    print("%d texts x %d lines, no mismatch" % (it, n))


if __name__ == "__main__":
    main(float(sys.argv[1]) if len(sys.argv) > 1 else 60.0, int(sys.argv[2]) if len(sys.argv) > 2 else 100000)
