"""Host-side data formats next to the path (no GPU): the `.smtx` reader (include/gunrock/io/smtx.hxx; reference
include/gunrock/io/smtx.hxx:56-169) and `gunrock::array` (include/gunrock/container/array.hxx; reference
include/gunrock/container/array.hxx:92-169).  Compiled as host code with nvcc (Thrust host vectors), nothing is
launched."""
import subprocess

from conftest import ROOT

SRC = r"""
#include <cstdio>
#include <iostream>
#include <gunrock/container/array.hxx>
#include <gunrock/io/smtx.hxx>
using namespace gunrock;
using namespace memory;
static_assert(sizeof(array<int, 3>) == 3 * sizeof(int), "aggregate, no overhead");
constexpr array<int, 3> kDims = {4, 5, 6};
static_assert(kDims.size() == 3 && kDims[1] == 5 && kDims.back() == 6 && !kDims.empty(), "constexpr access");
int main(int argc, char** argv) {
  array<float, 4> a;
  a.fill(2.5f);
  a[2] = 7.0f;
  float sum = 0;
  for (float x : a) sum += x;
  array<float, 4> b = a;
  std::printf("array %g %d %d %zu\n", sum, int(a == b), int(array<int, 0>{}.empty()), a.max_size());
  io::smtx_t<int, int, float> loader;
  try {
    auto csr = loader.load(argv[1], argc > 2);
    auto again = loader.load(argv[1], argc > 2);
    std::printf("smtx %s %d %d %d |", loader.dataset.c_str(), csr.number_of_rows, csr.number_of_columns,
                csr.number_of_nonzeros);
    for (auto x : csr.row_offsets) std::printf(" %d", x);
    std::printf(" |");
    for (auto x : csr.column_indices) std::printf(" %d", x);
    bool in_range = true, same = true;
    for (std::size_t i = 0; i < csr.nonzero_values.size(); ++i) {
      in_range = in_range && csr.nonzero_values[i] >= 1.0f && csr.nonzero_values[i] < 10.0f;
      same = same && csr.nonzero_values[i] == again.nonzero_values[i];
    }
    std::printf(" | %d %d\n", int(in_range), int(same));
  } catch (const std::invalid_argument& e) {
    std::printf("invalid_argument %s\n", e.what());
    return 3;
  } catch (const std::runtime_error& e) {
    std::printf("runtime_error %s\n", e.what());
    return 4;
  }
}
"""


def test_smtx_reader_and_array(tmp_path):
    src = tmp_path / "t.cu"
    src.write_text(SRC)
    exe = tmp_path / "t"
    subprocess.run(["nvcc", "-std=c++17", "--extended-lambda", "--expt-relaxed-constexpr",
                    "-gencode", "arch=compute_100a,code=sm_100a", f"-I{ROOT}/include", str(src), "-o", str(exe)],
                   check=True, timeout=600)

    def run(text, *extra, name="g.smtx"):
        p = tmp_path / name
        p.write_text(text)
        r = subprocess.run([str(exe), str(p), *extra], capture_output=True, text=True, timeout=60)
        return r.returncode, r.stdout.strip().splitlines()

    rc, out = run("% Sparse matrix file format .smtx\n%\n% a comment\n4 5 6\n0 2 2 5 6\n1 4 0 2 3 4\n")
    assert rc == 0 and out[0] == "array 14.5 1 1 4"
    assert out[1] == "smtx g 4 5 6 | 0 2 2 5 6 | 1 4 0 2 3 4 | 1 1"
    rc, out = run("3, 3, 2\n% comment between the data lines\n0 1 1 2\n2 0\n", "csv", name="h.smtx")
    assert rc == 0 and out[1] == "smtx h 3 3 2 | 0 1 1 2 | 2 0 | 1 1"
    rc, out = run("0 0 0\n0\n\n", name="empty.smtx")                 # no rows, no entries
    assert rc == 0 and out[1].startswith("smtx empty 0 0 0 | 0 |")
    # header / line-length disagreements and malformed content are rejected, as in the reference (smtx.hxx:148-166)
    assert run("4 5 6\n0 2 2 5\n1 4 0 2 3 4\n")[0] == 3              # one row offset short
    assert run("4 5 6\n0 2 2 5 6\n1 4 0 2 3\n")[0] == 3              # one column index short
    assert run("4 5 6\n0 2 1 5 6\n1 4 0 2 3 4\n")[0] == 3            # offsets decrease
    assert run("4 5 6\n0 2 2 5 6\n1 4 0 2 3 5\n")[0] == 3            # column id == columns
    assert run("4 5\n")[0] == 3                                      # truncated file
    r = subprocess.run([str(exe), str(tmp_path / "missing.smtx")], capture_output=True, text=True)
    assert r.returncode == 4 and "Unable to open file" in r.stdout
