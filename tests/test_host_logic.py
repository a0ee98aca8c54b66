"""Host-side logic shared by the tools (Python twin in vcfx_b200/api.py and shard.py): chunking,
header facts, allele_counter header/selection parsing, shard planning."""
import pytest

from vcfx_b200 import api, shard


def test_chunk_bounds_are_newline_aligned():
    data = b"".join(b"line%d\n" % i for i in range(1000)) + b"tail-without-newline"
    for cb in (32, 64, 1000, 10 ** 6):
        parts = list(api.chunk_bounds(data, cb))
        assert parts[0][0] == 0 and parts[-1][1] == len(data)
        for (s, e), (s2, _) in zip(parts, parts[1:]):
            assert e == s2 and data[e - 1:e] == b"\n" and e - s <= cb
    with pytest.raises(api.VcfxCudaError):
        list(api.chunk_bounds(b"x" * 100 + b"\n", 10))
    assert list(api.chunk_bounds(b"", 10)) == []


def test_header_facts():
    d = b"##a\n#CHROM\tPOS\n1\t2\n"
    assert api.find_chrom_header(d) == 4 and api.first_data_offset(d) == 15
    assert api.find_chrom_header(b"#CHROM\n") == 0
    assert api.find_chrom_header(b"1\t2\n") == 4 and api.first_data_offset(b"1\t2\n") == 0
    assert api.first_data_offset(b"##x\n##y") == 7
    assert api.first_data_offset(b"##x\n\n#y\n") == 4          # a blank line ends the block


def test_allele_counter_header_and_selection():
    d = b"##x\n#CHROM\tP\tI\tR\tA\tQ\tF\tI\tF\tS0\tS1\t\tS3\t\n#CHROM\tP\tI\tR\tA\tQ\tF\tI\tF\tT0\n1\t1\n"
    names, pos = api._ac_header_names(d)
    assert names == [b"S0", b"S1", b"", b"S3", b"T0"] and d[pos:] == b"1\t1\n"
    assert api._ac_parse_samples("A  B\t C") == [b"A", b"B", b"C"]
    assert api._ac_parse_samples(None) == []


def test_shard_plan_covers_the_data_and_keeps_lines_whole():
    hdr = b"##h\n#CHROM\tPOS\n"
    body = b"".join(b"1\t%d\t.\tA\tG\n" % i for i in range(997))
    data = hdr + body
    for world in (1, 2, 3, 8):
        plan = shard.plan(data, world)
        assert len(plan) == world and plan[0].start == 0 and plan[-1].end == len(data)
        line_no = 0
        for a, b in zip(plan, plan[1:]):
            assert a.end == b.start and data[a.end - 1:a.end] == b"\n"
        for s in plan:
            assert s.first_line == line_no + 1
            line_no += data[s.start:s.end].count(b"\n")
        assert [s.chrom_seen_before for s in plan] == [False] + [True] * (world - 1)
    # more ranks than lines: trailing shards are empty, nothing is lost
    plan = shard.plan(b"a\nb\n", 4)
    assert b"".join(b"a\nb\n"[s.start:s.end] for s in plan) == b"a\nb\n"
