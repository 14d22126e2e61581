"""include/cxxopts.hpp: our own stand-in for the cxxopts subset Gunrock's example programs use
(examples/algorithms/tc/tc.cu:23-45 is the pattern compiled here).  Host-only, g++."""
import subprocess

from conftest import ROOT

SRC = r"""
#include <cxxopts.hpp>
#include <iostream>
int main(int argc, char** argv) {
  cxxopts::Options options(argv[0], "Triangle Counting example");
  options.add_options()("help", "Print help")(
      "validate", "CPU validation", cxxopts::value<bool>()->default_value("false"))(
      "m,market", "Matrix file", cxxopts::value<std::string>())(
      "n,num_runs", "runs", cxxopts::value<int>()->default_value("1"))(
      "p,uniquify_percent", "percent", cxxopts::value<float>());
  try {
    auto result = options.parse(argc, argv);
    if (result.count("help") || (result.count("market") == 0)) {
      std::cout << options.help({""}) << std::endl;
      return 0;
    }
    std::cout << result["market"].as<std::string>() << "|" << result["validate"].as<bool>() << "|"
              << result["num_runs"].as<int>() << "|" << result.count("uniquify_percent");
    if (result.count("uniquify_percent") == 1) std::cout << "|" << result["uniquify_percent"].as<float>();
    std::cout << std::endl;
  } catch (const cxxopts::OptionException& e) {
    std::cout << "EXC " << e.what() << std::endl;
    return 3;
  }
}
"""


def test_cxxopts_subset(tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text(SRC)
    exe = tmp_path / "t"
    subprocess.run(["g++", "-std=c++17", f"-I{ROOT}/include", str(src), "-o", str(exe)], check=True)

    def run(*a):
        r = subprocess.run([str(exe), *a], capture_output=True, text=True)
        return r.returncode, r.stdout.strip()

    assert run("-m", "a.mtx") == (0, "a.mtx|0|1|0")
    assert run("--market=b.mtx", "--validate", "-n", "7", "--uniquify_percent", "12.5") == (0, "b.mtx|1|7|1|12.5")
    assert run("--market", "c.mtx", "--validate=false", "-n3") == (0, "c.mtx|0|3|0")
    rc, out = run("--help")
    assert rc == 0 and "--market arg" in out and "(default: false)" in out
    rc, out = run()
    assert rc == 0 and "Usage:" in out            # no --market: help, like the examples
    assert run("--nope")[0] == 3 and run("-m")[0] == 3


def test_reference_cmd_tool_compiles_unchanged_and_runs():
    """The reference's examples/tools/cmd.cu (its command-line parser test: <cxxopts.hpp> + util/filepath.hxx and
    nothing else) built UNCHANGED by examples/build_reference_examples.sh -> examples/bin/tool_cmd.  Host only."""
    import os
    import pytest
    exe = os.path.join(ROOT, "examples", "bin", "tool_cmd")
    if not os.path.exists(exe):
        pytest.skip("examples/bin/tool_cmd not built (needs /root/reference at build time)")

    def run(*a):
        r = subprocess.run([exe, *a], capture_output=True, text=True, timeout=30)
        return r.returncode, r.stdout

    assert run("-m", "graph.mtx") == (0, "")                 # a .mtx name: accepted silently
    assert run("--csr", "graph.csr", "-d", "1") == (0, "")
    rc, out = run("-m", "graph.txt")                          # not a market file: help
    assert rc == 0 and "Gunrock commandline parser test" in out and "--market arg" in out and "--csr arg" in out
    rc, out = run()
    assert rc == 0 and "Usage:" in out


def test_reference_unit_tests_compile_unchanged_and_the_host_only_ones_pass():
    """examples/bin/ref_unittests = the reference's own unit tests (/root/reference/unittests/unittests.hxx: formats,
    graph, memory, problem, parallel_for, advance merge_path, type limits, launch box, context, device properties,
    array, .smtx, triangle counting) compiled UNCHANGED against include/ with tests/gtest_shim standing in for
    googletest.  All 19 are in the binary; the ones that need no device run here."""
    import os
    import pytest
    exe = os.path.join(ROOT, "examples", "bin", "ref_unittests")
    if not os.path.exists(exe):
        pytest.skip("examples/bin/ref_unittests not built (needs /root/reference at build time)")
    listed = subprocess.run([exe, "--gtest_list_tests"], capture_output=True, text=True, timeout=30).stdout.split()
    assert len(listed) == 19 and {"graph.graph", "operators.prallel_for", "algorithm.tc", "io.smtx",
                                  "containers.array", "cuda.launch_box_occupancy"} <= set(listed)
    host_only = ["cuda.device_properties", "cuda.launch_box_fallback", "cuda.launch_box_define",
                 "operators_advance.merge_path_coordinate_struct", "operators_advance.merge_path_enum_exists"]
    r = subprocess.run([exe, "--gtest_filter=" + ":".join(host_only)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "5 tests ran, 0 failed" in r.stdout, r.stdout[-2000:] + r.stderr[-1000:]
    for name in host_only:
        assert f"[       OK ] {name}" in r.stdout
    # the fallback entry is what SM_TARGET = 100 selects in the reference's box (it lists sm_35 ... sm_86 only)
    assert "block_dimensions:    16, 1, 1" in r.stdout and "grid_dimensions:     2, 1, 1" in r.stdout
