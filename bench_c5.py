#!/usr/bin/env python
"""bench_c5.py — BASELINE config 5: a whole-genome-scale synthetic VCF (~50 GB of the C2 shape, 2,504 samples)
through VCFX_allele_freq_calc, resident in HBM.

  N = 1   the whole stream in one 50 GB device buffer, ONE launch over it (offsets far beyond 4 GiB, ~200 K tiles,
          ~5 M row records).  Parity: the output must equal, byte for byte, the concatenation of the outputs of the
          ~4.3 GB pieces the stream was generated in (each piece run on its own), and the last piece — the one that
          lies beyond 46 GB in the big buffer — must equal the CPU restatement (oracle/) of the same bytes.
  N > 1   (torchrun) strong scaling of that same stream: rank r owns the r-th newline-aligned range (whole lines),
          parses and reduces it on its own GPU; the row text is gathered on rank 0 in rank order (= file order) and
          its sha256 compared with the N = 1 run's (profiles/r2_c5.json).  What crosses GPUs: the gather of the
          finished text and one all-reduce of the totals; no data-path collective.

One JSON line on stdout (rank 0).  Not the contractual bench line — that is bench.py's.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
import bench as B  # noqa: E402

PIECE_VARIANTS = 427_409          # the C2 file, the unit the stream is generated in


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--pieces", type=int, default=12, help="number of C2-sized pieces in the whole stream (12 ~ 51.6 GB)")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    B.bind_to_gpu_numa_node(local_rank)
    import numpy as np
    import torch
    import torch.distributed as dist
    from vcfx_b200 import api, synth

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    threads = min(max(1, len(os.sched_getaffinity(0)) // max(world, 1)), 48)

    V_total = args.pieces * PIECE_VARIANTS
    v0 = V_total * rank // world
    v1 = V_total * (rank + 1) // world
    hdr = synth.header(2, B.SAMPLES, 2) if rank == 0 else b""
    bound = synth.line_bound(2, B.SAMPLES)
    # generate in C2-sized pieces through one pinned buffer; keep the piece boundaries (byte offsets in the device buffer)
    piece_cap = bound * PIECE_VARIANTS
    host = torch.empty(piece_cap, dtype=torch.uint8, pin_memory=True)
    hnp = host.numpy()
    # exact size first (the generator is deterministic): a dry pass would cost as much as the real one, so the device
    # buffer is sized from the bound of the average line (lines are ~10,065 B against a bound of ~10,100)
    d_in = torch.empty(len(hdr) + int((v1 - v0) * 10080) + (64 << 20) + api.DEVICE_PAD, dtype=torch.uint8, device=dev)
    if hdr:
        d_in[: len(hdr)].copy_(torch.frombuffer(bytearray(hdr), dtype=torch.uint8))
    t0 = time.perf_counter()
    off = len(hdr)
    cuts = [0]
    v = v0
    last_piece_host = None
    while v < v1:
        cnt = min(PIECE_VARIANTS, v1 - v)
        n = synth.lines_into(hnp, 2, B.SAMPLES, v, cnt, seed=2, threads=threads)
        assert off + n + api.DEVICE_PAD <= d_in.numel(), "device buffer too small for the stream"
        d_in[off:off + n].copy_(host[:n])
        off += n; v += cnt
        cuts.append(off)
        if v >= v1:
            last_piece_host = hnp[:n].copy() if (rank == 0 and world == 1) else None
    torch.cuda.synchronize()
    nbytes = off
    B.log(f"[c5 r{rank}] variants [{v0}, {v1}): {nbytes / 1e9:.2f} GB generated and uploaded in {time.perf_counter() - t0:.1f}s")
    del host

    out_cap = (v1 - v0) * 40 + (1 << 20)
    d_out = torch.empty(out_cap, dtype=torch.uint8, device=dev)
    ctx = api.Context(api.OP_ALLELE_FREQ, api.FILE, device=local_rank)
    ctx.set_line_hint(10065)
    vf = api.find_chrom_header(hdr) if rank == 0 else 0
    ms = []
    st = None
    for i in range(args.reps + 1):
        if world > 1:
            dist.barrier()
        ctx.run_device(d_in.data_ptr(), nbytes, d_out.data_ptr(), out_cap, valid_from=vf)
        st = ctx.sync()
        if i:
            ms.append(st.kernel_ms)
    assert st.rows == v1 - v0, (st.rows, v1 - v0)
    k_ms = sorted(ms)[len(ms) // 2]
    out_bytes = int(st.bytes_out)
    text = d_out[:out_bytes]

    peak = 6545.3
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        peak = float(json.loads(p.read_text())["hbm_gbs"])

    if world == 1:
        full_sha = hashlib.sha256(text.cpu().numpy().tobytes()).hexdigest()
        # the same stream piece by piece: every piece copied to an aligned scratch buffer and run on its own
        scratch = torch.empty(piece_cap + api.DEVICE_PAD, dtype=torch.uint8, device=dev)
        s_out = torch.empty(PIECE_VARIANTS * 40 + (1 << 20), dtype=torch.uint8, device=dev)
        h = hashlib.sha256()
        last_piece_out = b""
        for i in range(len(cuts) - 1):
            a = cuts[i] if i else 0          # piece 0 carries the header
            b = cuts[i + 1]
            scratch[: b - a].copy_(d_in[a:b])
            torch.cuda.synchronize()         # (the library launches on its own stream)
            ctx.run_device(scratch.data_ptr(), b - a, s_out.data_ptr(), s_out.numel(), valid_from=(vf if i == 0 else 0))
            s = ctx.sync()
            last_piece_out = s_out[: int(s.bytes_out)].cpu().numpy().tobytes()
            h.update(last_piece_out)
        pieces_sha = h.hexdigest()
        # the last piece against the CPU restatement
        from oracle import oracle as O

        class _S:            # the minimal shard interface oracle_parallel needs
            pass
        sh = _S()
        sh.hdr = synth.header(2, B.SAMPLES, 2)
        sh.hnp = np.concatenate([np.frombuffer(sh.hdr, dtype=np.uint8), last_piece_host])
        sh.nbytes = len(sh.hnp)
        exp_sha, exp_n = B.oracle_parallel(O, "allele_freq", sh, np)
        last_ok = hashlib.sha256(last_piece_out).hexdigest() == exp_sha and exp_n == len(last_piece_out)
        line = {"config": "C5", "tool": "allele_freq_calc", "n_gpus": 1, "variants": V_total, "samples": B.SAMPLES, "input_bytes": nbytes,
                "output_bytes": out_bytes, "kernel_ms": k_ms, "input_GBps": nbytes / k_ms / 1e6, "frac_of_hbm_peak": (nbytes + out_bytes) / k_ms / 1e6 / peak,
                "genotypes_per_s": V_total * B.SAMPLES / (k_ms / 1e3), "launches": "one launch over the whole 50 GB buffer",
                "parity": {"sha256": full_sha, "equals_concatenation_of_piece_runs": full_sha == pieces_sha,
                           "last_piece_equals_oracle": last_ok, "rows": int(st.rows), "rows_expected": V_total}}
        print(json.dumps(line))
    else:
        # strong scaling: gather the text on rank 0 in rank order, time = slowest rank
        t = torch.tensor([k_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([out_bytes], dtype=torch.int64, device=dev))
        sizes = [int(s.item()) for s in sizes]
        mx = max(sizes)
        padded = torch.zeros(mx, dtype=torch.uint8, device=dev)
        padded[:out_bytes].copy_(text)
        t0 = time.perf_counter()
        parts = [torch.empty(mx, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == 0 else None
        dist.gather(padded, parts, dst=0)
        torch.cuda.synchronize()
        gather_s = time.perf_counter() - t0
        tot = torch.tensor([nbytes, int(st.rows)], dtype=torch.int64, device=dev)
        dist.all_reduce(tot)
        if rank == 0:
            h = hashlib.sha256()
            for pt, n in zip(parts, sizes):
                h.update(pt[:n].cpu().numpy().tobytes())
            sha = h.hexdigest()
            ref = None
            rp = ROOT / "profiles" / "r2_c5.json"
            if rp.exists():
                try:
                    ref = json.loads(rp.read_text())["n1"]["parity"]["sha256"]
                except Exception:
                    ref = None
            ms_all = float(t.item())
            total_bytes = int(tot[0].item())
            line = {"config": "C5", "tool": "allele_freq_calc", "n_gpus": world, "scaling": "strong", "variants": V_total, "input_bytes": total_bytes,
                    "output_bytes": sum(sizes), "kernel_ms_slowest_rank": ms_all, "input_GBps": total_bytes / ms_all / 1e6,
                    "gather_seconds": gather_s, "rows": int(tot[1].item()),
                    "parity": {"sha256": sha, "n1_sha256": ref, "equal_to_one_gpu_run": (sha == ref) if ref else None, "rows_expected": V_total}}
            print(json.dumps(line))
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
