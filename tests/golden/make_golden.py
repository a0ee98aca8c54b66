#!/usr/bin/env python
"""Regenerate tests/golden/*.json from the compiled REFERENCE tools (oracle/_ref/VCFX_*, built by
`make -C oracle ref` from /root/reference).  Run in the build container only: the GPU box has no
reference tree, it just reads the committed fixtures.

Each fixture = one small input (base64) + the stdout / exit code of every reference tool on it, in
FILE mode (`-i file`) and STDIN mode.  Inputs are our own: the synthetic shapes of BASELINE.json at
toy size, seeded adversarial files, and hand-written cases that restate what the reference's test
scripts pin (tests/test_allele_freq_calc.sh, test_variant_counter.sh, test_hwe_tester.sh,
test_missing_detector.sh, test_allele_counter.sh — SURVEY.md §4)."""
import base64
import json
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from oracle import oracle as O          # noqa: E402
import vcfgen                           # noqa: E402
from vcfx_b200 import synth            # noqa: E402

H = "##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t"


def hand_cases():
    c = {}
    # allele frequencies 0.5 / 0.3333 / 0.8333 style cases, multi-allelic, phased, missing
    c["af_basic"] = (H + "S1\tS2\tS3\n"
                     "1\t100\trs1\tA\tG\t.\tPASS\t.\tGT\t0/1\t0/0\t1/1\n"
                     "1\t200\trs2\tA\tG,T\t.\tPASS\t.\tGT\t0/1\t0/2\t0/0\n"
                     "1\t300\trs3\tA\tG\t.\tPASS\t.\tGT:DP\t0|1:5\t.|.:0\t1|1:9\n"
                     "1\t400\trs4\tA\tG\t.\tPASS\t.\tGT\t1/1\t1/1\t0/1\n"
                     "1\t500\trs5\tA\tG\t.\tPASS\t.\tDP\t5\t6\t7\n")
    c["af_late_header"] = ("##fileformat=VCFv4.2\n1\t50\t.\tA\tG\t.\t.\t.\tGT\t0/1\n" + H.split("\n")[1] + "S1\n"
                           "1\t100\t.\tA\tG\t.\t.\t.\tGT\t0/1\n1\t200\t.\tA\n")
    # rounding tie: FILE prints 0.0313, STDIN 0.0312 (SURVEY.md finding 1)
    c["af_tie"] = H + "\t".join(f"S{i}" for i in range(16)) + "\n1\t1\t.\tA\tG\t.\t.\t.\tGT\t0/1" + "\t0/0" * 15 + "\n"
    c["vc_mixed"] = (H + "S1\n1\t1\t.\tA\tG\t.\tPASS\t.\tGT\t0/1\n1\t2\t.\tA\n\n#late comment\n1\t3\t.\tA\tG\t.\tPASS\tDP=1\n"
                     "1\t4\t.\tA\tG\t.\tPASS\n")
    c["hwe_basic"] = (H + "S1\tS2\tS3\tS4\n"
                      "1\t100\t.\tA\tG\t.\t.\t.\tGT\t0/0\t0/1\t1/1\t0/1\n"
                      "1\t200\t.\tA\tG,T\t.\t.\t.\tGT\t0/1\t1/2\t0/0\t0/0\n"
                      "1\t300\t.\tA\tG\t.\t.\t.\tDP:GT\t3:0/1\t3:0/1\t3:0/1\t3:0/1\n"
                      "1\t400\t.\tA\tG\t.\t.\t.\tGT\t0/1\t0/1\t0/1\t0/1\n"
                      "1\t500\t.\tA\tG\t.\t.\t.\tGT\t./.\t0|1\t1\t1/1\n")
    c["md_basic"] = (H + "S1\tS2\n"
                     "1\t100\t.\tA\tG\t.\tPASS\tDP=25\tGT\t0/1\t./.\n"
                     "1\t200\t.\tA\tG\t.\tPASS\t.\tGT\t.\t1\n"
                     "1\t300\t.\tA\tG\t.\tPASS\tAF=0.5;\tGT:GQ\t0/1:0.5\t.|1:3\n"
                     "1\t400\t.\tA\tG\t.\tPASS\tDP=1\tGT\t0/1\t1/1\n"
                     "1\t500\t.\tA\tG\n")
    c["ac_basic"] = (H + "X\tY\n"
                     "1\t100\trs1\tA\tG\t.\tPASS\t.\tGT\t0/1\t1/2\n"
                     "1\t200\trs2\tA\tG\t.\tPASS\t.\tGT\t./.\t0/.\n"
                     "1\t300\trs3\tA\tG\t.\tPASS\t.\tGT:DP\t0|0:3\t1|1:4\n"
                     "1\t400\n")
    # nonref_filter: what "definitely hom-ref" means in the two modes (000 vs 00, empty columns, GT not first, a final tab,
    # CRLF, a data line in front of the header, an unterminated last line)
    L = lambda fmt, *s: "1\t1\t.\tA\tG\t.\t.\t.\t" + fmt + "\t" + "\t".join(s)
    c["nr_quirks"] = ("1\t1\t.\tA\tG\t.\t.\t.\tGT\t0/0\n" + H + "S1\tS2\n" + "\n".join([
        L("GT", "0/0", "0|0"), L("GT", "0/0", "0|1"), L("GT", "000", "0/0"), L("GT", "00", "0"), L("GT", "//", "|"), L("GT", "0/0", ""),
        L("GT", "0/0") + "\t", L("GT"), "1\t1\t.\tA\tG\t.\t.\t.\tGT", L("GT:DP", "0/0:1", "0/0:"), L("DP:GT", "1:0/0", "1"),
        L("DP:GT", "1:0/0", "1:"), L("DP:GT", "1:0/0", ":0|0"), L("DP", "1", "2"), L("GT:", "0/0", "0/0"), L("", "0/0", "0/0"), "",
        L("GT", "0/0", "0/0\r"), L("GT", "0/0", "0/0") + "\r", "\r", L("GT", "0/0/0", "0|0|0|0"), L("GT", "./.", "0/0"), L("GT", "0/0", "0/0")]))
    # phase_checker: what "fully phased" means (three bytes x|y with x, y not '.'; otherwise every '|'-separated allele non-empty
    # and not "."), empty columns (dropped in the middle, ignored at the end of a file-mode line), GT not first, no GT key,
    # short lines, CRLF (content in stdin mode), a data line in front of the header, an unterminated last line
    c["pc_quirks"] = ("1\t1\t.\tA\tG\t.\t.\t.\tGT\t0|1\n" + H + "S1\tS2\n" + "\n".join([
        L("GT", "0|1", "1|0"), L("GT", "0|1", "0/1"), L("GT", "0|1", ".|."), L("GT", "0|1", "."), L("GT", "0|1", "0"), L("GT", "0|", "|1"),
        L("GT", "0|1|1", "10|2"), L("GT", "0|.|1", "0|1"), L("GT", "0||1", "0|1"), L("GT", "/|/", "||1"), L("GT", "..|1", ".1|0"), L("GT", "0|1", ""),
        L("GT", "0|1") + "\t", L("GT"), "1\t1\t.\tA\tG\t.\t.\t.\tGT", "1\t1\t.\tA\tG\t.\t.\t.\tDP", "1\t1\t.\tA\tG\t.\t.\t.\t", "1\t5\tx",
        L("GT:DP", "0|1:1", "1|1:"), L("DP:GT", "1:0|1", "1"), L("DP:GT", "1:0|1", "1:"), L("DP:GT", "1:0|1", ":1|0:7"), L("DP", "1", "2"),
        L("GT:", "0|1", "0|1"), L("", "0|1", "0|1"), "", L("GT", "0|1", "0|1\r"), L("GT", "0|1", "0|1") + "\r", "\r", "x\ty",
        L("GT", "1|2|3|4", "12|345"), L("GT", "./.", "0|1"), L("GT", "0|1", "0|1")]))
    # phase_checker, file mode: the FORMAT cache starts as ("", GT index 0) — an empty FORMAT column means "GT first" until the
    # first non-empty FORMAT was looked at (a line that ends behind its eighth tab does not get that far)
    c["pc_format_cache"] = (H + "S1\tS2\n" + "\n".join([
        L("", "0|1", "1|0"), L("", "0/1", "1|0"), "1\t1\t.\tA\tG\t.\t.\t.\t", "1\t2\tx", L("", "1|1:9", "2|1"), "", L("", "0|1") + "\t",
        L("DP", "1", "2"), L("", "0|1", "1|0"), L("GT", "0|1", "1|0"), L("", "0|1", "1|0")]) + "\n")
    # inbreeding_calculator: what a genotype code is (blanks / '\r' in front, nothing behind the second allele matters), lines
    # with fewer columns than samples (file mode keeps the code of the last line that had the column, stdin mode has none),
    # a final tab, multi-allelic and empty ALT, two "#CHROM" lines in the header block (file mode adds the names up), CRLF
    G = lambda alt, *s: "1\t7\t.\tA\t" + alt + "\t.\t.\t.\tGT\t" + "\t".join(s)
    c["ib_quirks"] = ("##fileformat=VCFv4.2\n" + H.split("\n")[1] + "S1\tS2\tS3\n\n#CHROM\tP\tI\tR\tA\tQ\tF\tI\tFORMAT\tS4\n" + "\n".join([
        G("G", "0/1", "0/0", "1/1", "0|1"), G("G", "1|0", " 0/1", "\r1/1", "0/1x"), G("G", "0/1", "1/1"), G("G,T", "0/1", "0/1", "0/1", "0/1"),
        G("", "0/1", "0/1", "0/1", "0/1"), G("G", "0/0", "0/1", "1/1", "0/0") + "\t", G("G", "0/1", "0/1", "0/1") + "\t", "#late\tline", "",
        G("G", "./.", ".", "0", "0/"), G("G", "00/01", "0/1/1", "0/2", "10/1"), G("G", "0/1:9", "1/1:3:4", "0:1/1", "0/1"), "1\t5\tx",
        G("<DEL>", "1/1", "1/1", "1/1", "1/1"), G("G", "0/0", "0/0", "0/0", "0/0"), G("G", "0/1", "0/0", "1/1", "0/1", "1/1", "0/0"),
        "1\t7\t.\tA\tG\t.\t.\t.\tGT\t", "1\t7\t.\tA\tG\t.\t.\t.\tGT", G("G", "0/1", "0/0", "1/1", "0/1\r"), G("G", "1/1", "0/1", "0/0", "0|0")]) + "\r\n")
    # genotype_query: flexible and strict matching (allele order, phasing, several digits, malformed genotypes), GT not the first
    # key / no GT key, lines with fewer than nine columns (with and without a final tab), empty lines, '#' lines between and
    # behind the data lines (stdin mode prints them only once a data line follows), '\r' is content
    c["gq_quirks"] = (H + "S1\tS2\n" + "\n".join([
        L("GT", "0/0", "0|1"), L("GT", "1|0", "0/0"), L("GT", "0/0", "1/1"), L("GT", "0/1/1", "0"), L("GT", ".", "./."), L("GT", "01/0", "0/01"),
        L("GT", "10/1", "1/10"), L("GT:DP", "0/0:3", "1/0:4"), L("DP:GT", "3:0/0", "4:0|1"), L("DP:GT", "3", "4:"), L("DP", "0/1", "0/1"), L("", "0/1", "0/1"),
        "#between\tdata", "", L("GT", "0/0", "") + "\t", L("GT", "", "0/1"), L("GT"), "1\t1\t.\tA\tG\t.\t.\t.\tGT", "1\t5\tx", "1\t5\t", "a\tb\tc\td\te\tf\tg\th",
        L("GT", "0/0", "0/1\r"), L("GT", "0/1", "0/0") + "\r", "\r", L("GT", "0|1:7", "0/0"), L("GT", "1/1", "0/0"), "#tail1", "", "#tail2"]) + "\n")
    # dosage_calculator: what a dosage is (separators anywhere, several digits, more or fewer than two alleles, a '.', stray bytes),
    # GT not the first key / no GT key, empty columns, a final tab (that column is not there), short lines, CRLF
    c["ds_quirks"] = (H + "S1\tS2\tS3\n" + "\n".join([
        L("GT", "0/0", "0|1", "1/1"), L("GT", "1|2", "10/0", "00/01"), L("GT", "/0/1", "0//1", "0/1/"), L("GT", "0", "0/1/2", "0/1x"), L("GT", ".", "./.", "1/."),
        L("GT", "", "0/1", ""), L("GT", "0/1", "1/1") + "\t", L("GT", "0/1") + "\t\t", L("GT"), "1\t1\t.\tA\tG\t.\t.\t.\tGT", "1\t5\tx", "",
        L("GT:DP", "0/1:3", "1/1:", "0/0"), L("DP:GT", "3:0/1", "4", "5:"), L("DP:GT", ":1/1", "::", "3:1|0:9"), L("DP", "0/1", "1/1", "0/0"), L("", "0/1", "1/1", "0/0"),
        "#late\tline", L("GT", " 0/1", "0/1 ", "0:1"), L("GT", "0/1", "1/1", "0/1\r"), L("GT", "1/1", "0/0", "0/1") + "\r", "\r", L("GT", "2|3", "0|0", "1|0")]) + "\n")
    # indexer: what counts as CHROM and POS in the two modes (blanks in front, signs, wrap-around, CRLF, "#CHROM" look-alikes)
    c["ix_quirks"] = ("##x\n #CHROMX\tY\n#CHROM\tPOS\tID\n1\t100\t.\n 2\t+7x\t.\n\t3\t5\n4\t0\n5\t-3\n6\t 12\n7\n8\t99999999999999999999\n"
                      "9\t12\r\n\n#late\t1\nchrX\t007\tid\n10\t9223372036854775807\n11\t9223372036854775808\n12\t-9223372036854775808\n13\t5")
    return {k: v.encode() for k, v in c.items()}


def run_all(data: bytes, ac_ok: bool, md_file_ok: bool = True):
    out = {}
    with tempfile.NamedTemporaryFile(suffix=".vcf") as f:
        f.write(data); f.flush()
        for tool in ("allele_freq_calc", "hwe_tester", "missing_detector"):
            extra = ["-t", "1"] if tool == "missing_detector" else []
            if not (tool == "missing_detector" and not md_file_ok):
                rc, so, _ = O.run_ref(tool, ["-q", *extra, "-i", f.name])
                out[f"{tool}.file"] = [rc, base64.b64encode(so).decode()]
            rc, so, _ = O.run_ref(tool, ["-q"] if tool != "hwe_tester" else [], stdin=data)
            out[f"{tool}.stdin"] = [rc, base64.b64encode(so).decode()]
        rc, so, se = O.run_ref("indexer", [f.name])
        out["indexer.file"] = [rc, base64.b64encode(so).decode(), se.count(b"no #CHROM")]
        rc, so, se = O.run_ref("indexer", [], stdin=data)
        out["indexer.stdin"] = [rc, base64.b64encode(so).decode(), se.count(b"no #CHROM")]
        rc, so, se = O.run_ref("nonref_filter", ["-i", f.name])
        out["nonref_filter.file"] = [rc, base64.b64encode(so).decode(), se.count(b"Warning")]
        rc, so, se = O.run_ref("nonref_filter", [], stdin=data)
        out["nonref_filter.stdin"] = [rc, base64.b64encode(so).decode(), se.count(b"Warning")]
        # phase_checker: stdout and the whole stderr text ("-" = stdin without the reference's own "is there anything on
        # stdin yet" probe, which shows the help text when the pipe is not filled in time)
        rc, so, se = O.run_ref("phase_checker", ["-i", f.name])
        out["phase_checker.file"] = [rc, base64.b64encode(so).decode(), base64.b64encode(se).decode()]
        rc, so, se = O.run_ref("phase_checker", ["-"], stdin=data)
        out["phase_checker.stdin"] = [rc, base64.b64encode(so).decode(), base64.b64encode(se).decode()]
        # dosage_calculator: both modes, stdout and the whole stderr text
        for key, args, stdin in (("file", ["-i", f.name], None), ("stdin", [], data)):
            rc, so, se = O.run_ref("dosage_calculator", args, stdin=stdin)
            out[f"dosage_calculator.{key}"] = [rc, base64.b64encode(so).decode(), base64.b64encode(se).decode()]
        # genotype_query: a flexible and a strict query in both modes, stdout and the whole stderr text
        for key, args, stdin in (("file.het", ["-g", "0/1", "-i", f.name], None), ("stdin.het", ["-g", "1/0"], data),
                                 ("file.strict", ["-g", "0|1", "--strict", "-i", f.name], None), ("stdin.strict", ["-g", "1/1", "--strict"], data)):
            rc, so, se = O.run_ref("genotype_query", args, stdin=stdin)
            out[f"genotype_query.{key}"] = [rc, base64.b64encode(so).decode(), base64.b64encode(se).decode()]
        # inbreeding_calculator: default options in both modes, and the other frequency options
        for key, args, stdin in (("file", ["-q", "-i", f.name], None), ("stdin", ["-q"], data), ("file.global", ["-q", "--freq-mode", "global", "-i", f.name], None),
                                 ("file.skipcount", ["-q", "--skip-boundary", "--count-boundary-as-used", "-i", f.name], None),
                                 ("stdin.skip", ["-q", "--skip-boundary"], data)):
            rc, so, se = O.run_ref("inbreeding_calculator", args, stdin=stdin)
            out[f"inbreeding_calculator.{key}"] = [rc, base64.b64encode(so).decode(), base64.b64encode(se).decode()]
        for strict in (False, True):
            a = ["--strict"] if strict else []
            rc, so, se = O.run_ref("variant_counter", [*a, f.name])
            out[f"variant_counter.file.strict{int(strict)}"] = [rc, base64.b64encode(so).decode(), se.count(b"Warning")]
            rc, so, se = O.run_ref("variant_counter", a, stdin=data)
            out[f"variant_counter.stdin.strict{int(strict)}"] = [rc, base64.b64encode(so).decode(), se.count(b"Warning")]
        if ac_ok:
            for key, args, stdin in (("mt", ["-q", "-i", f.name], None), ("stream", ["-q"], data), ("agg", ["-q", "-a", "-i", f.name], None),
                                     ("bin", ["-q", "-b", "-i", f.name], None), ("limit2", ["-q", "-l", "2", "-i", f.name], None)):
                rc, so, _ = O.run_ref("allele_counter", args, stdin=stdin)
                out[f"allele_counter.{key}"] = [rc, base64.b64encode(so).decode()]
    return out


def main():
    assert O.have_reference(), "build the reference tools first: make -C oracle ref"
    fixtures = {}
    for name, data in hand_cases().items():
        # (the reference allele_counter does not come back from these two inputs: no allele_counter outputs for them)
        fixtures[name] = dict(input=base64.b64encode(data).decode(), expect=run_all(data, ac_ok=(name not in ("ib_quirks", "gq_quirks", "ds_quirks"))))
    for shape, V, S in ((1, 40, 12), (2, 12, 300), (3, 30, 60), (4, 12, 9)):
        data = synth.make_vcf(shape, V, S, seed=70 + shape)
        fixtures[f"shape{shape}"] = dict(input=base64.b64encode(data).decode(), expect=run_all(data, ac_ok=True))
    for seed in range(6):
        data = vcfgen.make_vcf(7000 + seed, n_lines=25, n_samples=2 + seed, crlf=(seed == 2), final_newline=(seed != 3),
                               header=["normal", "late", "normal", "none", "double", "normal"][seed])
        fixtures[f"fuzz{seed}"] = dict(input=base64.b64encode(data).decode(), expect=run_all(data, ac_ok=False))
        data = vcfgen.make_vcf(7100 + seed, n_lines=25, n_samples=2 + seed, domain="ac", final_newline=(seed != 3),
                               header=["normal", "late", "normal", "none", "double", "normal"][seed])
        fixtures[f"acfuzz{seed}"] = dict(input=base64.b64encode(data).decode(), expect=run_all(data, ac_ok=True))
    (HERE / "reference_outputs.json").write_text(json.dumps(fixtures, indent=0, sort_keys=True))
    print(f"wrote {len(fixtures)} fixtures, {(HERE / 'reference_outputs.json').stat().st_size} bytes")


if __name__ == "__main__":
    main()
