"""The N>1 path on CPU: two ranks over gloo.  Each rank takes its newline-aligned shard from
vcfx_b200.shard.plan and runs it through the library's streaming entry points — the product's own kernel and
C-ABI sources under the warp emulator (tests/emu/), with the shard's prefix facts (header seen before the
shard, end of the leading '#' block, first line number) passed exactly as a GPU rank passes them; the totals
cross ranks in one all-reduce and the texts are concatenated in rank order.  The result must equal the oracle
on the whole file."""
import os
import socket
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, data, q):
    sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
    import torch
    import torch.distributed as dist
    from oracle import oracle as O
    from vcfx_b200 import api, shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, str(ROOT / "tests" / "emu"))
    import build_emu
    dev = build_emu.load_api()                             # libvcfx_emu.so: the kernels, emulated
    sh = shard.plan(data, world)[rank]
    body = data[sh.start:sh.end]
    chunk = 64 << 10

    def run(op, mode, valid_abs=0, flags=0):
        ctx = dev.Context(op, mode, flags=flags, chunk_bytes=chunk)
        outs, t = dev.stream_bytes(ctx, body, chunk, valid_abs)
        ctx.close()
        return b"".join(outs), t
    # prefix facts: data lines in front of the "#CHROM" line do not count for allele_freq_calc; the leading '#'
    # block bounds missing_detector's pre-scan; line numbers continue from the shards before
    af_rows, af_t = run(dev.OP_ALLELE_FREQ, dev.FILE, shard.valid_from(sh, dev.find_chrom_header(data)))
    _, vc_t = run(dev.OP_VARIANT_COUNT, dev.FILE)
    bad = (sh.first_line - 1 + vc_t.first_short_line) if vc_t.first_short_line else 1 << 62
    md_out, md_t = run(dev.OP_MISSING_DETECT, dev.STDIN, min(max(sh.header_block_end - sh.start, 0), len(body)))

    class _R:
        pass
    af = _R(); af.rows = af_t.rows
    md = _R(); md.flagged = md_t.flagged; md.out = md_out
    vc_rows = vc_t.rows
    tot = torch.tensor([af.rows, vc_rows, md.flagged], dtype=torch.int64)
    dist.all_reduce(tot)                                   # the path's only exchange: scalar totals
    first_bad = torch.tensor([bad], dtype=torch.int64)
    dist.all_reduce(first_bad, op=dist.ReduceOp.MIN)       # --strict: the smallest failing line wins
    parts = [None] * world
    dist.all_gather_object(parts, (af_rows, md.out))
    if rank == 0:
        q.put((api.AF_HEADER + b"".join(p[0] for p in parts), b"".join(p[1] for p in parts), tot.tolist(), int(first_bad.item())))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_two_ranks_equal_one(world, oracle):
    import torch.multiprocessing as mp
    sys.path.insert(0, str(ROOT / "tests"))
    import vcfgen
    from vcfx_b200 import synth
    cases = [synth.make_vcf(3, 400, 50, seed=21),
             vcfgen.make_vcf(31, n_lines=300, n_samples=4, header="late"),
             vcfgen.make_vcf(32, n_lines=5, n_samples=2)]
    ctx = mp.get_context("spawn")
    for data in cases:
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, world, port, data, q)) for r in range(world)]
        for p in procs: p.start()
        af_out, md_out, tot, first_bad = q.get(timeout=120)
        for p in procs: p.join(timeout=60)
        assert all(p.exitcode == 0 for p in procs)
        O = oracle
        assert af_out == O.allele_freq(data, O.FILE).out
        assert md_out == O.missing(data, O.STDIN).out
        assert tot[0] == O.allele_freq(data, O.FILE).rows and tot[1] == O.variant_count(data).rows
        strict = O.variant_count(data, O.FILE, True)
        assert (first_bad if strict.first_bad_line else 0) == strict.first_bad_line
