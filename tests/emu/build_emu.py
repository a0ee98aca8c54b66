"""Builds tests/emu/_build/libvcfx_emu.so: the product's kernel + C-ABI sources compiled with g++ against
the warp emulator in this directory.  TEST INFRASTRUCTURE ONLY — nothing under vcfx_b200/ knows this exists."""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
OUT = HERE / "_build" / "libvcfx_emu.so"


def build(force: bool = False) -> Path:
    srcs = [ROOT / "vcfx_b200" / "csrc" / "vcfx_api.cu", HERE / "cuda_emu.cpp"]
    deps = srcs + sorted((ROOT / "vcfx_b200" / "csrc").glob("*.cuh")) + sorted((ROOT / "include").glob("*.h")) + [HERE / "cuda_runtime.h"]
    if not force and OUT.exists() and all(d.stat().st_mtime <= OUT.stat().st_mtime for d in deps):
        return OUT
    OUT.parent.mkdir(exist_ok=True)
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-ffp-contract=off", "-Wall", "-Wno-unused-function", "-Wno-unknown-pragmas", "-Wno-unused-variable", "-x", "c++", "-DVCFX_EMU", "-DVCFX_C4_RING=1",
           "-I", str(HERE), "-I", str(ROOT / "include"), "-I", str(ROOT / "vcfx_b200" / "csrc"),
           "-fPIC", "-shared", "-o", str(OUT), *map(str, srcs)]
    tmp = OUT.with_name(f".{OUT.name}.{os.getpid()}")               # (several test processes may build at once: link aside, rename)
    cmd[cmd.index("-o") + 1] = str(tmp)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("emulator build failed")
    os.replace(tmp, OUT)
    return OUT


def build_tools(force: bool = False) -> Path:
    """The five drop-in executables linked against the emulator library (tests/emu/_build/bin): the host side of the
    tools — chunking, carry, prefix facts, ordered output, several contexts in one process — runs for real, the
    kernels under the emulator."""
    so = build(force)
    tools_dir = ROOT / "vcfx_b200" / "tools"
    out_dir = HERE / "_build" / "bin"
    out_dir.mkdir(parents=True, exist_ok=True)
    common = sorted(p for p in tools_dir.glob("*.cpp") if not p.name.startswith("VCFX_"))
    hdrs = sorted(tools_dir.glob("*.h")) + sorted((ROOT / "include").glob("*.h"))
    for src in sorted(tools_dir.glob("VCFX_*.cpp")):
        exe = out_dir / src.stem
        deps = [src, *common, *hdrs, so]
        if not force and exe.exists() and all(d.stat().st_mtime <= exe.stat().st_mtime for d in deps):
            continue
        tmp = exe.with_name(f".{exe.name}.{os.getpid()}")          # (several test processes may build at once: link aside, rename)
        cmd = ["g++", "-O1", "-g", "-std=c++17", "-I", str(ROOT / "include"), "-I", str(tools_dir), str(src), *map(str, common),
               "-o", str(tmp), f"-L{so.parent}", "-lvcfx_emu", f"-Wl,-rpath,{so.parent}", "-lz", "-lpthread", "-ldl", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            raise RuntimeError("emulator tool build failed")
        os.replace(tmp, exe)
    return out_dir


def load_api():
    """A private copy of vcfx_b200.api bound to the emulator library (the real module stays untouched)."""
    import importlib.util
    so = build()
    spec = importlib.util.spec_from_file_location("vcfx_b200._api_under_emulator", ROOT / "vcfx_b200" / "api.py")
    mod = importlib.util.module_from_spec(spec)
    mod.__package__ = "vcfx_b200"
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    mod.lib_path = lambda: so
    mod.load()
    return mod


if __name__ == "__main__":
    print(build(force=True))
