#!/usr/bin/env python
"""Which memory path do the visited-bit probes of push-BFS level 1 take, default kernel vs the on-chip-copy variants?
(design input, host only: the kernels run under the CPU emulator of tests/cuemu)

Builds the bench graph family at a reduced scale with the oracle generator, writes it in `.csr` layout, compiles
tests/cuemu/emu_profile.cpp against the kernel section of include/gunrock/b200/advance.cuh (as tests/test_cuemu_kernels.py
does) and prints its table.  The copy is sized in proportion to the graph as the launcher sizes it for scale 22
(1.27 M of 4.19 M vertices per CTA; 2x / 4x for the cluster variants), the level is shared by 16 CTAs, each cluster with
its own copy taken at kernel start.  Usage: python profiles/micro/emulated_probe_paths.py [scale] > profiles/r1_emulated_probe_paths.txt"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402


def main():
    scale = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    import types
    conftest = types.ModuleType("conftest")
    conftest.ROOT = os.path.abspath(ROOT)
    sys.modules["conftest"] = conftest
    import test_cuemu_kernels as t
    ro, ci = oracle.rmat_csr(scale, 16, 0x5EED22)
    src = int(np.diff(ro).argmax())
    with tempfile.TemporaryDirectory() as d:
        adv = open(os.path.join(ROOT, "include", "gunrock", "b200", "advance.cuh")).read()
        cut = adv.rindex("// ----", 0, adv.index("// Host launchers"))
        open(os.path.join(d, "advance_kernels.gen.cuh"), "w").write(adv[:cut] + "\n}\n}\n")
        bfs = open(os.path.join(ROOT, "include", "gunrock", "b200", "bfs.cuh")).read()
        open(os.path.join(d, "functors.gen.cuh"), "w").write(
            "#pragma once\nnamespace gunrock {\nnamespace b200 {\n" + t.block_from(bfs, "struct bfs_claim_op {") + "}\n}\n")
        exe = os.path.join(d, "emu_profile")
        subprocess.check_call(["g++", "-std=c++20", "-O2", "-pthread", "-Wno-attributes", "-I", os.path.join(ROOT, "tests", "cuemu"),
                               "-I", d, "-I", os.path.join(ROOT, "include"), "-o", exe,
                               os.path.join(ROOT, "tests", "cuemu", "emu_profile.cpp")])
        path = os.path.join(d, "g.csr")
        with open(path, "wb") as f:
            np.array([len(ro) - 1, len(ro) - 1, len(ci)], np.int32).tofile(f)
            ro.tofile(f)
            ci.tofile(f)
        print(f"RMAT scale {scale}, ef 16, seed 0x5EED22 (bench graph family)")
        sys.stdout.flush()
        subprocess.check_call([exe, path, str(src)])


if __name__ == "__main__":
    main()
