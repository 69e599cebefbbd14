#!/usr/bin/env python3
"""Fixture recipe (test infrastructure, not product code): the 13-read crafted known-answer test.

Writes kat.fa and kat.sam into the directory given as argv[1].  Every read was designed to hit one
quirk of the reference `junc` path (see DESIGN.md "Quirks"): entropy loop, secondary / placed-unmapped
records, junction-wide window walking through a neighbouring intron, single-end reads counted as R2,
nb_dist_aln order dependence, I/D next to splice sites, soft clips, read N vs genome N, soft-masking,
cross-target distance sentinels.
"""
import sys, os

def genome(L, salt):
    s = []
    for i in range(L):
        h = ((i + salt) * 2654435761) & 0xFFFFFFFF
        h ^= h >> 15; h = (h * 2246822519) & 0xFFFFFFFF; h ^= h >> 13
        s.append("ACGT"[h & 3])
    return s

def build():
    A = genome(3000, 0); B = genome(2000, 7777)
    def put(g, pos, s):
        for k, c in enumerate(s): g[pos + k] = c
    put(A, 200, "GT"); put(A, 298, "AG")      # intron I1 [200,299]  GT..AG  (+)
    put(A, 400, "CT"); put(A, 498, "AC")      # intron I2 [400,499]  CT..AC  (-)
    put(A, 260, "NNNN")                       # Ns inside I1
    put(A, 330, "N")                          # N inside exon 2 (read N vs genome N = match)
    for k in range(340, 360): A[k] = A[k].lower()   # soft-masked exon region
    put(B, 1000, "GC"); put(B, 1098, "AG")    # intron on chrB [1000,1099] GC..AG semi-canonical (+)
    fa = []
    for name, g in (("chrA", A), ("chrB", B)):
        fa.append(">%s\n" % name)
        for i in range(0, len(g), 60): fa.append("".join(g[i:i+60]) + "\n")
    GA = "".join(A).upper(); GB = "".join(B).upper()
    def sub(s, idx):  # substitute base at idx with the next base in ACGT
        c = s[idx]; n = "ACGT"[("ACGT".find(c) + 1) % 4] if c in "ACGT" else "A"
        return s[:idx] + n + s[idx+1:]
    reads = []
    def add(name, flag, ref, pos, mapq, cigar, seq, extra="", mref="*", mpos=0, tlen=0):
        reads.append((0 if ref == "chrA" else 1, pos, "%s\t%d\t%s\t%d\t%d\t%s\t%s\t%d\t%d\t%s\t*%s" % (
            name, flag, ref, pos+1, mapq, cigar, mref, mpos, tlen, seq, ("\t"+extra) if extra else "")))
    add("r01", 99, "chrA", 150, 60, "50M100N50M", GA[150:200] + GA[300:350], "XS:A:+", "=", 401, 400)
    add("r02", 99, "chrA", 150, 60, "50M100N60M", GA[150:200] + GA[300:360], "XS:A:+", "=", 401, 400)
    add("r03", 147, "chrA", 150, 60, "50M100N50M", sub(GA[150:200], 47) + GA[300:350], "XS:A:+", "=", 101, -400)
    add("r04", 99, "chrA", 160, 60, "40M100N100M100N20M", GA[160:200] + GA[300:400] + GA[500:520], "XS:A:+", "=", 601, 500)
    add("r05", 0, "chrA", 180, 60, "20M100N250M", GA[180:200] + GA[300:550], "")
    add("r06", 16, "chrA", 170, 3, "28M2I2M100N10M3D37M", GA[170:198] + "TT" + GA[198:200] + GA[300:310] + GA[313:350], "XS:A:+")
    add("r07", 99, "chrA", 185, 60, "5S15M100N30M5S", "ACGTA" + GA[185:200] + GA[300:330] + "TTTTT", "XS:A:-", "=", 451, 400)
    add("r08", 355, "chrA", 310, 0, "90M100N30M", GA[310:400] + GA[500:530], "XS:A:-", "=", 701, 500)
    add("r09", 99, "chrA", 320, 60, "80M100N40M", sub(sub(GA[320:400], 79), 10) + sub(GA[500:540], 0), "XS:A:-", "=", 701, 500)
    add("r10", 99, "chrA", 700, 60, "50M", GA[700:750], "", "=", 901, 300)
    add("r11", 4, "chrA", 800, 0, "*", GA[800:830], "")
    add("r12", 99, "chrB", 950, 60, "50M100N50M", GB[950:1000] + GB[1100:1150], "XS:A:+", "=", 1301, 400)
    add("r13", 163, "chrB", 960, 60, "40M100N60M", GB[960:1000] + GB[1100:1160], "XS:A:+", "=", 1301, 400)
    reads.sort(key=lambda r: (r[0], r[1]))   # stable: r01,r02,r03,r04,r06,r05,r07,r08,r09,r10,r11 | r12,r13
    sam = ["@HD\tVN:1.0\tSO:coordinate\n@SQ\tSN:chrA\tLN:3000\n@SQ\tSN:chrB\tLN:2000\n"]
    for _, _, l in reads: sam.append(l + "\n")
    return "".join(fa), "".join(sam)

if __name__ == "__main__":
    out = sys.argv[1] if len(sys.argv) > 1 else "."
    os.makedirs(out, exist_ok=True)
    fa, sam = build()
    open(os.path.join(out, "kat.fa"), "w").write(fa)
    open(os.path.join(out, "kat.sam"), "w").write(sam)
