#!/usr/bin/env python
"""bench.py — throughput of the VCFX hot path on B200 (BASELINE.json metric, config C2).

One "step" = the job of BASELINE config 2: VCFX_allele_freq_calc + VCFX_variant_counter over
one synthetic 1000G-chr21-shape VCF (427,409 variants x 2,504 samples, phased GT-only,
~4.3 GB) — each tool makes its own full pass, exactly like the two reference processes.

  value   input GB/s of that job with the file already resident in HBM (kernel-only)
  e2e     the same job through the C ABI from HOST memory (vcfx_cuda_submit_host /
          next_output): every step re-uploads the file in pinned 64 MiB chunks and reads the
          text back, H2D and D2H inside the timed region
  --impl reference : the unmodified reference tools (oracle/_ref/VCFX_*, built by
          oracle/Makefile from the reference sources) on this box's host cores, each step on
          a bounded sample of the same workload

N > 1 (torchrun, one rank per GPU): every rank owns its own newline-aligned shard of the same
synthetic stream (weak scaling, no data-path collective); the scalar totals cross ranks in one
tiny NCCL all-reduce per step.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

C2_VARIANTS = 427_409
C2_SAMPLES = 2_504
SHAPE = 2
SEED = 2
CHUNK = 64 << 20
CPU_SAMPLE_VARIANTS = 60_000          # ~0.6 GB: ~3-4 s of allele_freq_calc on one core


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, indices):
        """indices: physical GPU indices to watch (rank 0 watches every GPU of the job: one NVML poller
        per node instead of one per rank, so the polling itself does not perturb the launches)."""
        super().__init__(daemon=True)
        self.indices = list(indices)
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.hs = [pynvml.nvmlDeviceGetHandleByIndex(i) for i in self.indices]
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.hs[0], pynvml.NVML_CLOCK_SM) if self.hs else None
            self.ok = bool(self.hs)
        except Exception as e:  # pragma: no cover
            log(f"[bench] NVML unavailable: {e}")

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._halt.is_set():
            for h in self.hs:
                try:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                    try:
                        r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                    except Exception:
                        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for bit, name in names.items():
                        if r & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
            time.sleep(0.003)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_mhz_min": min(self.samples) if self.samples else None, "gpus_watched": len(getattr(self, "hs", [])),
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------
def physical_gpu_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


def cpu_reference_run(sample_variants: int, steps: int, warmup: int):
    """Time the reference's own CPU tools on a bounded sample of the workload (rank 0 only).

    Method = the reference's benchmark harness (benchmarks/scripts/run_comprehensive_benchmark.sh:
    122-146): warm page cache, `tool -i file > /dev/null`, wall clock."""
    from vcfx_b200 import synth
    ref_dir = ROOT / "oracle" / "_ref"
    af, vc = ref_dir / "VCFX_allele_freq_calc", ref_dir / "VCFX_variant_counter"
    if not (af.exists() and vc.exists()):
        return None
    data = synth.make_vcf(SHAPE, sample_variants, C2_SAMPLES, seed=SEED)
    d = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    fd, path = tempfile.mkstemp(suffix=".vcf", dir=d)
    times = []
    try:
        with os.fdopen(fd, "wb") as f:
            f.write(data)
        with open(path, "rb") as f:        # warm the page cache
            while f.read(1 << 24):
                pass
        devnull = open(os.devnull, "wb")
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            subprocess.run([str(af), "-q", "-i", path], stdout=devnull, stderr=devnull, check=True)
            subprocess.run([str(vc), path], stdout=devnull, stderr=devnull, check=True)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    finally:
        os.unlink(path)
    nbytes = len(data)
    return {"bytes": nbytes, "variants": sample_variants, "times": times,
            "gbps": nbytes / (sum(times) / len(times)) / 1e9}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    r = cpu_reference_run(CPU_SAMPLE_VARIANTS, args.steps, max(1, min(args.warmup, 1)))
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/VCFX_* not built (run make -C oracle ref where /root/reference exists)"}))
        return 0
    ms = 1e3 * sum(r["times"]) / len(r["times"])
    sample = f"{r['variants']} variants x {C2_SAMPLES} samples ({r['bytes'] / 1e9:.2f} GB) of the C2 stream, file in page cache"
    line = {
        "impl": "reference", "metric": "vcf_input_GB_per_s", "value": r["gbps"], "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(),
        "genotypes_per_s": r["variants"] * C2_SAMPLES / (ms / 1e3),
        "cpu_baseline": {"value": r["gbps"], "unit": "GB/s", "cores": 1, "kind": "reference", "sample": sample},
        "e2e": {"value": r["gbps"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def workload_config():
    return {"workload": "C2: VCFX_allele_freq_calc + VCFX_variant_counter (FILE semantics) on a synthetic "
                        "1000G chr21-shape VCF, 427409 variants x 2504 samples, phased GT-only, ~4.3 GB per GPU",
            "variants_per_gpu": C2_VARIANTS, "samples": C2_SAMPLES, "chunk_bytes": CHUNK,
            "l2": "inputs (4.3 GB) are larger than L2 (126 MB); no explicit flush",
            "sharding": "one newline-aligned shard of the stream per GPU; the scalar totals of the timed jobs are merged by one NCCL all-reduce inside the timed region"}


# ----------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="vcfx_b200", choices=["vcfx_b200", "reference"])
    ap.add_argument("--variants", type=int, default=C2_VARIANTS, help="variants per GPU (default: full C2)")
    ap.add_argument("--tile", type=int, default=0, help="tile_bytes for the resident path (0 = library default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    from vcfx_b200 import api, synth

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libvcfx_cuda has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    V = args.variants
    warmup = max(args.warmup, 3)

    # ---- this rank's shard of the stream, generated straight into pinned host memory
    t0 = time.perf_counter()
    hdr = synth.header(SHAPE, C2_SAMPLES, SEED) if rank == 0 else b""
    cap = len(hdr) + synth.line_bound(SHAPE, C2_SAMPLES) * V
    host = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
    hnp = host.numpy()
    hnp[: len(hdr)] = np.frombuffer(hdr, dtype=np.uint8)
    threads = max(1, (os.cpu_count() or 8) // max(world, 1))
    nbody = synth.lines_into(hnp[len(hdr):], SHAPE, C2_SAMPLES, rank * V, V, seed=SEED, threads=min(threads, 48))
    nbytes = len(hdr) + nbody
    log(f"[bench r{rank}] generated {nbytes / 1e9:.3f} GB ({V} variants) in {time.perf_counter() - t0:.1f}s")

    d_in = torch.empty(nbytes + api.DEVICE_PAD, dtype=torch.uint8, device=dev)
    d_in[:nbytes].copy_(host[:nbytes], non_blocking=False)
    torch.cuda.synchronize()
    out_cap = 64 << 20
    d_out = torch.empty(out_cap, dtype=torch.uint8, device=dev)
    # a real (non-default) stream: the library launches on it and torch's events time it
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    ctx_af = api.Context(api.OP_ALLELE_FREQ, api.FILE, device=local_rank, stream=stream, chunk_bytes=CHUNK, tile_bytes=args.tile)
    ctx_vc = api.Context(api.OP_VARIANT_COUNT, api.FILE, device=local_rank, stream=stream, chunk_bytes=CHUNK, tile_bytes=args.tile)
    valid_from = 0 if rank else api.find_chrom_header(hdr)
    totals = torch.zeros(2, dtype=torch.int64, device=dev)

    def step_resident():
        ctx_af.run_device(d_in.data_ptr(), nbytes, d_out.data_ptr(), out_cap, valid_from=valid_from)
        ctx_vc.run_device(d_in.data_ptr(), nbytes, 0, 0)

    def reduce_totals(n_jobs: int):
        """The path's only exchange: the scalar totals of the jobs just run, merged with ONE all-reduce
        (a [jobs x 2] int64 tensor) inside the timed region: the shards never talk to each other while they
        are being parsed, so no rank waits for another one between two jobs."""
        if world > 1:
            dist.all_reduce(totals.repeat(max(n_jobs, 1), 1))

    # ---- correctness gate (size-independent properties at full size)
    step_resident()
    torch.cuda.synchronize()
    st_af, st_vc = ctx_af.sync(), ctx_vc.sync()
    assert st_af.rows == V and st_vc.rows == V, (st_af.rows, st_vc.rows, V)
    out_bytes_af = int(st_af.bytes_out)
    head = bytes(d_out[:64].cpu().numpy())
    assert head.startswith(b"21\t"), head
    totals[0] = st_af.rows; totals[1] = st_vc.rows

    for _ in range(warmup):
        step_resident()
    reduce_totals(warmup)
    torch.cuda.synchronize()
    # NVML is set up BEFORE the barrier: done after it, the few milliseconds it takes on rank 0 made rank 0
    # enter the timed region late and every other rank wait for it at the closing all-reduce (measured:
    # +0.4 ms/step at 2 GPUs, +0.7 at 8)
    sampler = ClockSampler([physical_gpu_index(r) for r in range(world)] if rank == 0 else [])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    e0.record()
    t_host0 = time.perf_counter()
    for _ in range(args.steps):
        step_resident()
    reduce_totals(args.steps)
    e1.record()
    host_ms_per_step = (time.perf_counter() - t_host0) * 1e3 / args.steps      # enqueue time only (no sync yet)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    st_af, st_vc = ctx_af.sync(), ctx_vc.sync()
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    per_rank_ms = [ms_total / args.steps]
    if world > 1:                      # every rank's own time, for the record (value uses the slowest)
        g = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(g, torch.tensor([ms_total / args.steps], dtype=torch.float64, device=dev))
        per_rank_ms = [round(float(x.item()), 4) for x in g]
    log(f"[bench r{rank}] timed region: {ms_total / args.steps:.3f} ms/step on the device, {host_ms_per_step:.3f} ms/step of host enqueue")
    value = world * nbytes / (ms_step / 1e3) / 1e9

    # per-kernel durations (CUDA events around each tool's kernels on the launch stream)
    af_ms, vc_ms = [], []
    for _ in range(5):
        ctx_af.run_device(d_in.data_ptr(), nbytes, d_out.data_ptr(), out_cap, valid_from=valid_from)
        af_ms.append(ctx_af.sync().kernel_ms)
        ctx_vc.run_device(d_in.data_ptr(), nbytes, 0, 0)
        vc_ms.append(ctx_vc.sync().kernel_ms)
    af_k, vc_k = statistics.median(af_ms), statistics.median(vc_ms)
    log(f"[bench r{rank}] kernels alone: allele_freq_calc {af_k:.3f} ms, variant_counter {vc_k:.3f} ms")

    # ---- e2e: host buffers through the streaming C ABI, H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        ctx_af_s = api.Context(api.OP_ALLELE_FREQ, api.FILE, device=local_rank, chunk_bytes=CHUNK, n_slots=3)
        ctx_vc_s = api.Context(api.OP_VARIANT_COUNT, api.FILE, device=local_rank, chunk_bytes=CHUNK, n_slots=3)
        bounds = []
        pos = 0
        while pos < nbytes:
            end = min(nbytes, pos + CHUNK)
            if end < nbytes:
                end = pos + int(np.flatnonzero(hnp[pos:end] == 10)[-1]) + 1
            bounds.append((pos, end)); pos = end
        base_ptr = host.data_ptr()

        def step_e2e():
            """The job through the C ABI from host memory: every chunk is uploaded ONCE
            (vcfx_cuda_submit_host on the allele_freq_calc context) and variant_counter runs on the same
            device bytes (vcfx_cuda_submit_shared); both texts come back to the host."""
            rows_af = rows_vc = nout = 0

            def drain_pair():
                nonlocal rows_af, rows_vc, nout
                out, st, _ = ctx_vc_s.next_output(); rows_vc += st.rows; nout += len(out)
                out, st, _ = ctx_af_s.next_output(); rows_af += st.rows; nout += len(out)

            for (s, e) in bounds:
                vf = min(max(valid_from - s, 0), e - s)
                while not ctx_af_s.submit_host(base_ptr + s, e - s, valid_from=vf, is_final=(e == nbytes)):
                    drain_pair()
                ok = ctx_vc_s.submit_shared(ctx_af_s, is_final=(e == nbytes))
                assert ok
            while ctx_af_s.in_flight():
                drain_pair()
            assert rows_af == V and rows_vc == V, (rows_af, rows_vc)
            if world > 1:
                dist.all_reduce(totals)
            return nout

        for _ in range(2):
            d2h = step_e2e()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e2e_steps = max(3, min(args.steps, 5))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            d2h = step_e2e()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item()) / e2e_steps
        e2e = {"value": world * nbytes / e2e_s / 1e9, "unit": "GB/s", "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
               "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": int(d2h),
               "note": "pinned host memory -> 64 MiB chunks, 3 slots in flight; each chunk is uploaded once and feeds both tools (vcfx_cuda_submit_host + vcfx_cuda_submit_shared)"}
        ctx_af_s.close(); ctx_vc_s.close()

    # ---- roofline of the dominant kernel
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak = float(json.loads(peaks_path.read_text())["hbm_gbs"]); peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak = 6650.0; peak_src = "B200_PROFILING.md fallback (of fallback)"
    alg_bytes = nbytes + out_bytes_af
    achieved = alg_bytes / (af_k / 1e3) / 1e9
    # DRAM traffic of the dominant kernel from the committed ncu --set full capture of this workload
    traffic = None
    prof = ROOT / "profiles" / "r1_ncu_full_c2.json"
    if prof.exists() and V == C2_VARIANTS:
        try:
            k = json.loads(prof.read_text())[0]
            traffic = (float(k["dram__bytes_read.sum"].split()[0]) + float(k["dram__bytes_write.sum"].split()[0]) / 1e3) * 1e9
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "vcfx_scan_kernel<OP_AF> (+ tile_scan + format_rows: the tool's kernels)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "algorithmic_bytes_per_launch": alg_bytes,
                "kernel_ms": af_k, "peak_source": peak_src,
                "other_kernels": {"vcfx_scan_kernel<OP_VC>+resolve_events": {"kernel_ms": vc_k,
                                  "achieved": nbytes / (vc_k / 1e3) / 1e9, "frac": nbytes / (vc_k / 1e3) / 1e9 / peak}}}

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:      # reported at N=1 only
            r = cpu_reference_run(CPU_SAMPLE_VARIANTS, 2, 1)
            if r:
                cpu = {"value": r["gbps"], "unit": "GB/s", "cores": 1, "kind": "reference",
                       "sample": f"{r['variants']} variants x {C2_SAMPLES} samples ({r['bytes'] / 1e9:.2f} GB) of the C2 stream, "
                                 f"allele_freq_calc -q -i + variant_counter, best-effort warm cache, {os.cpu_count()} host cores present"}
        line = {
            "metric": "vcf_input_GB_per_s", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": workload_config(),
            "genotypes_per_s": world * V * C2_SAMPLES / (ms_step / 1e3),
            "variants_per_s": world * V / (ms_step / 1e3),
            "bytes_per_gpu": nbytes,
            "e2e": e2e, "gpu_launches": 5 * args.steps, "per_rank_ms_per_step": per_rank_ms, "clocks": clocks,
            "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    ctx_af.close(); ctx_vc.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
