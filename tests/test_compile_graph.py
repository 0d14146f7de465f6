"""Host-only checks of the graph compiler (csrc/pus_graph.hpp -> the HBM image) through tools/compile_graph_hash.cpp: random graphs
with loop closures, duplicate observations and removed factors compile, a second compile into the same object gives the same
image (the harness folds that into `ok`), and the image is a pure function of the graph (same hash from a second process).
The harness is what proves host-side refactors of compile_graph bit-identical (tools/README.md)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("cgh") / "cgh")
    hdr = os.path.join(ROOT, "pop_up_slam_b200", "csrc", "pus_graph.hpp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I" + os.path.dirname(hdr), '-DHDR="%s"' % hdr, "-x", "c++",
                           os.path.join(ROOT, "tools", "compile_graph_hash.cpp"), "-o", out])
    return out


@pytest.mark.parametrize("args", [("17", "5", "11"), ("300", "14", "2"), ("2500", "16", "15"), ("700", "8", "18", "3"), ("6000", "12", "16")])
def test_compiled_image_is_deterministic(harness, args):
    runs = [subprocess.run([harness, *args], capture_output=True, text=True, check=True).stdout for _ in range(2)]
    m = [re.search(r"ok=(\d) E=(\d+) hash=([0-9a-f]{16})", r) for r in runs]
    assert all(m), runs
    assert m[0].group(1) == "1", runs[0]           # compiled, and the re-compile into the same object succeeded
    assert int(m[0].group(2)) > 0
    assert m[0].group(3) == m[1].group(3)          # same graph -> same image
