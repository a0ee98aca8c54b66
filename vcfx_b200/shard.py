"""Multi-GPU decomposition: newline-aligned byte ranges, one per rank — the reference's own
thread-chunking rule (allele_counter.cpp:890-903, missing_detector.cpp:404-422) lifted to GPUs.

Records are independent, so the data path needs no collective: rank r parses, reduces and formats
its range, the outputs are concatenated in rank order, and only prefix facts cross ranks — whether a
"#CHROM" line lies before the shard (allele_freq_calc.cpp:372-386), the 1-based number of its first
line (variant_counter warnings) — plus the scalar totals, which are one tiny all-reduce.
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass
class Shard:
    rank: int
    start: int
    end: int
    first_line: int            # 1-based line number of the shard's first line
    chrom_seen_before: bool    # a line starting with "#CHROM" lies before `start`
    header_block_end: int      # absolute offset where the leading '#' block ends


def plan(data, world: int) -> list[Shard]:
    """Cut `data` (bytes-like) into `world` contiguous ranges that end on '\\n' (the last one may not)."""
    n = len(data)
    cuts = [0]
    for r in range(1, world):
        pos = max(cuts[-1], (n * r) // world)
        if pos < n:
            nl = data.find(b"\n", pos)
            pos = n if nl < 0 else nl + 1
        cuts.append(min(pos, n))
    cuts.append(n)
    from . import api
    chrom = api.find_chrom_header(data)
    hb_end = api.first_data_offset(data)
    out, line = [], 1
    for r in range(world):
        s, e = cuts[r], cuts[r + 1]
        out.append(Shard(r, s, e, line, chrom < s, hb_end))
        line += bytes(data[s:e]).count(b"\n")
    return out


def valid_from(sh: Shard, chrom_offset: int) -> int:
    """allele_freq_calc: offset inside the shard below which data lines precede the header."""
    return min(max(chrom_offset - sh.start, 0), sh.end - sh.start)
