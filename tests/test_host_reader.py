"""The tools' byte mover (vcfx_b200/tools/vcfx_host.cpp) without a GPU: Source::read with its
multi-threaded pread path for regular files, write_all with its multi-threaded pwrite path, pipes
falling back to read(2)/write(2).  The chunking above it (newline cut, carry) needs the library and a
device and is covered by tests/test_gpu_cli.py."""
import os
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
TOOLS = ROOT / "vcfx_b200" / "tools"


@pytest.fixture(scope="module")
def probe(tmp_path_factory):
    from vcfx_b200 import build
    lib = build.build_cuda()                       # vcfx_host.cpp links against the C ABI (never called here)
    exe = tmp_path_factory.mktemp("probe") / "reader_probe"
    cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-I", str(ROOT / "include"), "-I", str(TOOLS),
           str(ROOT / "tests" / "cpp" / "reader_probe.cpp"), str(TOOLS / "vcfx_host.cpp"),
           "-o", str(exe), f"-L{lib.parent}", "-lvcfx_cuda", f"-Wl,-rpath,{lib.parent}", "-lz", "-lpthread", "-ldl", "-lrt"]
    subprocess.run(cmd, check=True, capture_output=True)
    return exe


def _pattern(n: int) -> bytes:
    blk = bytes((i * 131 + (i >> 8) * 17) & 0xFF for i in range(1 << 16))
    return (blk * (n // len(blk) + 1))[:n]


@pytest.mark.parametrize("size", [0, 1, (8 << 20) - 1, 8 << 20, (8 << 20) + 1, (21 << 20) + 12345, (32 << 20) + 5])
@pytest.mark.parametrize("threads", ["1", "8"])
def test_copy_through_source_and_writer(probe, tmp_path, size, threads):
    src = tmp_path / "in.bin"; dst = tmp_path / "out.bin"
    data = _pattern(size)
    src.write_bytes(data)
    env = dict(os.environ, VCFX_IO_THREADS=threads)
    for cap in (16 << 20, (9 << 20) + 7):          # slices, and a cap that is not a multiple of anything
        r = subprocess.run([str(probe), str(src), str(dst), str(cap)], capture_output=True, env=env, timeout=120)
        assert r.returncode == 0, r.stderr
        assert int(r.stdout) == size
        assert dst.read_bytes() == data, (size, threads, cap)


@pytest.mark.parametrize("threads", ["1", "8"])
def test_reads_with_a_carry_keep_one_file_offset(probe, tmp_path, threads):
    """A chunk that starts with a carried-over tail asks for less than the pread threshold (read(2)), the next
    one for a full buffer (pread): both must follow the same file offset (the cached offset
    went stale and 42 MB came out as 75 MB)."""
    data = _pattern((41 << 20) + 4321)
    src = tmp_path / "in.bin"; dst = tmp_path / "out.bin"
    src.write_bytes(data)
    env = dict(os.environ, VCFX_IO_THREADS=threads)
    for cap, carry in ((8 << 20, 1 << 20), (8 << 20, 12345), ((9 << 20) + 7, (1 << 20) + 3)):
        r = subprocess.run([str(probe), str(src), str(dst), str(cap), str(carry)], capture_output=True, env=env, timeout=120)
        assert r.returncode == 0, r.stderr
        assert int(r.stdout) == len(data)
        assert dst.read_bytes() == data, (threads, cap, carry)


def test_pipes_use_plain_read_and_write(probe, tmp_path):
    data = _pattern((10 << 20) + 99)
    src = tmp_path / "in.bin"; src.write_bytes(data)
    # stdin from a pipe, stdout to a pipe: /dev/stdin and /dev/stdout of the child
    p1 = subprocess.Popen(["cat", str(src)], stdout=subprocess.PIPE)
    r = subprocess.run([str(probe), "/dev/stdin", "/dev/stdout", str(12 << 20)], stdin=p1.stdout, capture_output=True, timeout=120,
                       env=dict(os.environ, VCFX_IO_THREADS="8"))
    p1.wait()
    assert r.returncode == 0, r.stderr
    out = r.stdout
    tail = (b"%d\n" % len(data))
    assert out.endswith(tail) and out[: -len(tail)] == data
