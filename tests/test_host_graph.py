"""CPU test of the threaded host half of the graph phase (link-record sample sort, vertex grouping, GFA text in chunks; host_graph.cpp,
host_par.h): the text built on all host threads must equal the text of the sequential code path, on a synthetic graph large enough for
the threaded path. (Its content against the reference is the GPU suite's job: goldens, 1 M-read SHA-256.)"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "spades_b200")
SRC = os.path.join(ROOT, "tests", "host", "host_graph_check.cpp")
BIN = os.path.join(ROOT, "tests", "host", "_build", "host_graph_check")


def _build():
    lib = os.path.join(LIBDIR, "libspades_b200.so")
    if not os.path.exists(lib):
        pytest.skip("libspades_b200.so not built")
    if os.path.exists(BIN) and os.path.getmtime(BIN) > max(os.path.getmtime(SRC), os.path.getmtime(lib)):
        return
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    subprocess.check_call(["nvcc", "-std=c++17", "-O2", "-x", "cu", SRC, "-o", BIN, "-L" + LIBDIR, "-lspades_b200", "-Xlinker", "-rpath," + LIBDIR],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def _run(threads):
    env = dict(os.environ)
    if threads:
        env["SGPU_HOST_THREADS"] = str(threads)
    else:
        env.pop("SGPU_HOST_THREADS", None)
    out = subprocess.run([BIN, "150000"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-500:]
    return out.stdout.split()


def test_threaded_gfa_text_equals_sequential():
    _build()
    seq = _run(1)
    par = _run(0)
    three = _run(3)
    assert seq[4] == par[4] == three[4] == "0"                       # the sort primitive against std::sort
    assert seq[3] == "1"
    if (os.cpu_count() or 1) > 1:
        assert int(par[3]) > 1, "the threaded path was not exercised"
    assert seq[:3] == par[:3] == three[:3]                            # hash, bytes, lines of the GFA text
    assert int(seq[2]) > 150000                                       # one S line per edge plus links
