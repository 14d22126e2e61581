"""The barrier / value slots shared by the host threads of a multi-device run (include/gunrock/b200/thread_hub.hxx,
used by part_loops.cuh thread_exchange_t): lock-step, the one-barrier reduction pattern, abort instead of dead-lock."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hub_binary(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hub") / "hub_selftest")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-pthread", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "hub", "hub_selftest.cpp"), "-o", out])
    return out


@pytest.mark.parametrize("threads", [2, 3, 8, 16])
def test_thread_hub_barrier_reduction_and_abort(hub_binary, threads):
    r = subprocess.run([hub_binary, str(threads), "1500"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout + r.stderr
