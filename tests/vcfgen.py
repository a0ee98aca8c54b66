"""Small adversarial VCF generator for the parity tests (pure Python, seeded).

Covers the edge cases the reference's own tests and SURVEY.md Appendix B list: data before
``#CHROM``, '#' and blank lines between records, truncated and ragged lines, empty fields,
trailing tabs, CRLF, a missing final newline, FORMAT strings with GT first / later / absent /
as a prefix of another key, and genotype spellings (phased, unphased, missing, half-missing,
haploid, polyploid, multi-digit, leading zeros, empty).

``domain="ac"`` restricts the bytes of every sample's first ':' piece to ``[0-9./|]`` and
forbids CR, because the reference allele_counter loops forever outside that alphabet
(allele_counter.cpp:272-293) — there is nothing to compare against there.
"""
from __future__ import annotations

import random

GT_COMMON = ["0/0", "0/1", "1/1", "0|0", "0|1", "1|0", "1|1"]
GT_ODD = [".", "./.", ".|.", "./1", "0/.", "1|.", ".|0", "0", "1", "2", "1/2", "2|3", "0/2", "10/0", "0/10",
          "00/1", "01/1", "0/1/1", "0|0|1", "1/", "/1", "", "0//1", "12", "0/0/0/0", "./0/1", "3/3",
          "0.", ".0", "0./1", "1/.0"]
GT_NON_AC = [" 0/0", "0/1 ", "a/b", "0/x", "N", "0-1", "0\\1", "./A", "1/1\r"]   # outside [0-9./|]
FORMATS = ["GT", "GT", "GT", "GT:DP", "GT:AD:DP:GQ:PL", "DP:GT", "AD:DP", "GTX", "G", "", "GT:GT", "XGT:GT", "DP:GQ:GT"]
INFOS = [".", "", "DP=3", "DP=3;", "AF=0.5", "AC=1;AN=2", "NS=3;DP=14;AF=0.5;DB;H2"]
CHROMS = ["1", "21", "chr1", "X", "", "chrUn_gl000220"]
IDS = [".", "rs123", "", "rs1;rs2"]
REFS = ["A", "C", "G", "T", "AT", "", "N"]
ALTS = ["A", "C", "G", "T", "AT", "G,T", "A,C,G", "", ".", "<DEL>", "T,"]


def _sample(rng: random.Random, fmt: str, domain: str, p_odd: float) -> str:
    n_keys = max(1, fmt.count(":") + 1)
    pieces = []
    for k in range(n_keys):
        is_gt_like = (k == 0) or (fmt.split(":")[k:k + 1] == ["GT"])
        if is_gt_like:
            x = rng.random()
            if x < p_odd:
                g = rng.choice(GT_ODD)
            elif x < p_odd * 1.3 and domain != "ac":
                g = rng.choice(GT_NON_AC)
            else:
                g = rng.choice(GT_COMMON)
            if domain == "ac" and k == 0:
                g = "".join(c for c in g if c in "0123456789./|")
            pieces.append(g)
        else:
            if domain == "ac" and k == 0:
                pieces.append(str(rng.randrange(0, 60)))
            else:
                pieces.append(rng.choice(["35", "0", ".", "1,2", "10,0,255", "0.5", "99"]))
    # ragged: sometimes drop trailing pieces or add extras
    x = rng.random()
    if x < 0.05 and len(pieces) > 1:
        pieces = pieces[: rng.randrange(1, len(pieces))]
    elif x < 0.08:
        pieces.append("7")
    return ":".join(pieces)


def make_line(rng: random.Random, n_samples: int, domain: str = "any", p_odd: float = 0.25,
              p_bad_line: float = 0.15) -> str:
    fmt = rng.choice(FORMATS)
    fields = [rng.choice(CHROMS) if rng.random() < 0.2 else "1",
              str(rng.randrange(1, 10 ** 6)) if rng.random() > 0.03 else "",
              rng.choice(IDS), rng.choice(REFS) if rng.random() < 0.3 else "A",
              rng.choice(ALTS) if rng.random() < 0.3 else "G",
              rng.choice(["100", ".", "29.5", ""]), rng.choice(["PASS", ".", "q10"]),
              rng.choice(INFOS), fmt]
    ns = n_samples
    x = rng.random()
    if x < p_bad_line * 0.3:
        ns = rng.randrange(0, n_samples + 1)                  # fewer sample columns
    elif x < p_bad_line * 0.4:
        ns = n_samples + rng.randrange(1, 3)                  # extra columns
    fields += [_sample(rng, fmt, domain, p_odd) for _ in range(ns)]
    x = rng.random()
    if x < p_bad_line * 0.25:
        fields = fields[: rng.randrange(1, 10)]               # truncated line
    line = "\t".join(fields)
    x = rng.random()
    if x < p_bad_line * 0.15:
        line += "\t"                                          # trailing tab
    elif x < p_bad_line * 0.2 and domain != "ac":
        line = "\t" + line                                   # shifts FORMAT into the sample columns
    return line


def make_vcf(seed: int, n_lines: int = 40, n_samples: int = 5, domain: str = "any",
             crlf: bool = False, final_newline: bool = True, p_odd: float = 0.25,
             header: str = "normal", noise: bool = True) -> bytes:
    """``header``: 'normal' | 'none' | 'late' (a data line precedes #CHROM) | 'double'."""
    rng = random.Random(seed)
    names = [f"S{i}" for i in range(n_samples)]
    chrom_line = "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT" + "".join("\t" + n for n in names)
    out = ["##fileformat=VCFv4.2", "##source=vcfgen"]
    if header == "late":
        out.append(make_line(rng, n_samples, domain, p_odd))
    if header != "none":
        out.append(chrom_line)
    if header == "double":
        out.append(chrom_line)
    for _ in range(n_lines):
        x = rng.random()
        if noise and x < 0.03:
            out.append("")
        elif noise and x < 0.06:
            out.append("#a comment\twith\ttabs\t.\t.\t.\t.\t.\t.\t./.\t.")
        elif noise and x < 0.08 and domain != "ac":
            out.append("\r" if crlf else " ")
        else:
            out.append(make_line(rng, n_samples, domain, p_odd))
    eol = "\r\n" if crlf else "\n"
    text = eol.join(out)
    if final_newline:
        text += eol
    return text.encode("latin-1")
