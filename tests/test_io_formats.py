"""Host-side data formats next to the path (no GPU): the `.smtx` reader (include/gunrock/io/smtx.hxx; reference
include/gunrock/io/smtx.hxx:56-169) and `gunrock::array` (include/gunrock/container/array.hxx; reference
include/gunrock/container/array.hxx:92-169).  Compiled as host code with nvcc (Thrust host vectors), nothing is
launched."""
import os
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


SMTX_DUMP_SRC = r"""
#include <cstdio>
#include <gunrock/io/smtx.hxx>
using namespace gunrock;
int main(int argc, char** argv) {
  io::smtx_t<int, int, float> loader;
  auto csr = loader.load(argv[1]);
  std::printf("%d %d %d\n", csr.number_of_rows, csr.number_of_columns, csr.number_of_nonzeros);
  FILE* f = fopen(argv[2], "wb");
  fwrite(csr.row_offsets.data(), 4, csr.row_offsets.size(), f);
  fwrite(csr.column_indices.data(), 4, csr.column_indices.size(), f);
  fclose(f);
}
"""


def test_smtx_reader_against_the_reference_dataset_golden(tmp_path):
    """Golden vector: the reference's one `.smtx` dataset read by the reference's own `io::smtx_t::load`
    (tests/golden/make_golden.py -> swin_tiny_attn_proj_smtx.npz); its unit test asserts 96 x 96 with 4608 entries
    (/root/reference/unittests/io/smtx.cuh:28-30).  Our reader must give the same structure on the same file --
    re-serialised from the golden arrays, and the original where the reference tree is present."""
    import numpy as np

    g = np.load(os.path.join(ROOT, "tests", "golden", "swin_tiny_attn_proj_smtx.npz"))
    rows, cols, nnz = (int(x) for x in g["shape"])
    assert (rows, cols, nnz) == (96, 96, 4608)
    src = tmp_path / "s.cu"
    src.write_text(SMTX_DUMP_SRC)
    exe = tmp_path / "s"
    subprocess.run(["nvcc", "-std=c++17", "--extended-lambda", "--expt-relaxed-constexpr",
                    "-gencode", "arch=compute_100a,code=sm_100a", f"-I{ROOT}/include", str(src), "-o", str(exe)],
                   check=True, timeout=600)
    again = tmp_path / "again.smtx"
    again.write_text("% Sparse matrix file format .smtx\n% re-serialised from tests/golden\n"
                     f"{rows}, {cols}, {nnz}\n".replace(", ", " ")
                     + " ".join(str(int(x)) for x in g["row_offsets"]) + "\n"
                     + " ".join(str(int(x)) for x in g["column_indices"]) + "\n")
    files = [str(again)]
    original = "/root/reference/datasets/layers.0.blocks.0.attn.proj_swin_tiny_unstructured_50.smtx"
    if os.path.exists(original):
        files.append(original)
    for path in files:
        out = tmp_path / "dump.bin"
        r = subprocess.run([str(exe), path, str(out)], capture_output=True, text=True, timeout=60)
        assert r.returncode == 0 and r.stdout.split() == ["96", "96", "4608"], (path, r.stdout, r.stderr)
        raw = np.fromfile(out, dtype=np.int32)
        assert np.array_equal(raw[:rows + 1], g["row_offsets"]) and np.array_equal(raw[rows + 1:], g["column_indices"])


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


MTX_SRC = r"""
#include <cstdio>
#include <gunrock/io/matrix_market.hxx>
using namespace gunrock;
int main(int argc, char** argv) {
  io::matrix_market_t<int, int, float> mm;
  auto [props, coo] = mm.load(argv[1]);
  std::printf("%d %d %d w%d s%d d%d\n", coo.number_of_rows, coo.number_of_columns, coo.number_of_nonzeros,
              int(props.weighted), int(props.symmetric), int(props.directed));
  FILE* f = fopen(argv[2], "wb");
  const std::size_t n = coo.row_indices.size();
  fwrite(coo.row_indices.data(), 4, n, f);
  fwrite(coo.column_indices.data(), 4, n, f);
  fwrite(coo.nonzero_values.data(), 4, n, f);
  fclose(f);
}
"""


def test_matrix_market_loader_parallel_fast_path_equals_the_entry_loop(tmp_path):
    """include/gunrock/io/matrix_market.hxx parses a clean body with all host threads and leaves every irregular
    body to the entry-at-a-time fscanf loop (the reference's reading order, matrix_market.hxx:153-250).  Whatever
    the thread count, the outcome -- arrays bit for bit, properties, or the failure -- must be that loop's; clean
    files are also compared with the unmodified reference loader (oracle/_ref) and the C restatement."""
    import numpy as np
    import oracle

    src = tmp_path / "m.cu"
    src.write_text(MTX_SRC)
    exe = tmp_path / "m"
    subprocess.run(["nvcc", "-std=c++17", "-O1", "--extended-lambda", "--expt-relaxed-constexpr",
                    "-gencode", "arch=compute_100a,code=sm_100a", f"-I{ROOT}/include", str(src), "-o", str(exe)],
                   check=True, timeout=600)
    rng = np.random.default_rng(7)
    n = 5000
    r, c = rng.integers(1, 301, n), rng.integers(1, 301, n)
    w = rng.standard_normal(n) * 1e3
    body_real = "\n".join(f"{a} {b} {float(x)!r}" for a, b, x in zip(r, c, w))
    files = {
        "real_general": "%%MatrixMarket matrix coordinate real general\n% a comment\n%\n300 300 5000\n" + body_real + "\n",
        "no_final_newline": "%%MatrixMarket matrix coordinate real general\n300 300 5000\n" + body_real,
        "pattern_symmetric": "%%MatrixMarket matrix coordinate pattern symmetric\n300 300 5000\n"
                             + "\n".join(f"{max(a, b)} {min(a, b)}" for a, b in zip(r, c)) + "\n",
        "integer_general": "%%MatrixMarket MATRIX Coordinate Integer General\n300 300 5000\n"
                           + "\n".join(f"{a} {b} {int(x)}" for a, b, x in zip(r, c, w)) + "\n",
        "blanks_tabs_crlf": "%%MatrixMarket matrix coordinate real general\n9 9 4\n\n  1\t2   3.5  \r\n\n3 4 -1e-3\r\n"
                            "5 6 1E5\n   \n7 8 .25\n\n",
        "odd_numbers": "%%MatrixMarket matrix coordinate real general\n9 9 4\n1 2 inf\n3 4 0x1p-3\n5 6 1e-50\n7 8 -0.0\n",
        "empty_body": "%%MatrixMarket matrix coordinate pattern general\n5 5 0\n",
        # ---- irregular bodies: the fast path must step aside --------------------------------------------------
        "entry_split_over_lines": "%%MatrixMarket matrix coordinate real general\n9 9 3\n1 2\n3.5\n3 4 1\n5\n6 2\n",
        "two_entries_per_line": "%%MatrixMarket matrix coordinate pattern general\n9 9 4\n1 2 3 4\n5 6 7 8\n",
        "more_entries_than_announced": "%%MatrixMarket matrix coordinate pattern general\n9 9 2\n1 2\n3 4\n5 6\n7 8\n",
        "plus_sign": "%%MatrixMarket matrix coordinate pattern general\n9 9 2\n+1 2\n3 4\n",
        "short_file": "%%MatrixMarket matrix coordinate pattern general\n9 9 3\n1 2\n3 4\n",
        "zero_index": "%%MatrixMarket matrix coordinate pattern general\n9 9 2\n1 2\n0 4\n",
        "comment_in_body": "%%MatrixMarket matrix coordinate pattern general\n9 9 2\n1 2\n% no\n3 4\n",
        "value_missing": "%%MatrixMarket matrix coordinate real general\n9 9 2\n1 2 1.5\n3 4\n",
        "garbage_value": "%%MatrixMarket matrix coordinate real general\n9 9 2\n1 2 1.5x\n3 4 2\n",
    }
    clean = {"real_general", "no_final_newline", "pattern_symmetric", "integer_general", "blanks_tabs_crlf",
             "odd_numbers", "empty_body"}
    for name, text in files.items():
        path = tmp_path / f"{name}.mtx"
        path.write_text(text, newline="")
        outcomes = {}
        for threads in ("1", "2", "3", "7", None):
            env = dict(os.environ)
            env.pop("GUNROCK_B200_MTX_THREADS", None)
            if threads:
                env["GUNROCK_B200_MTX_THREADS"] = threads
            out = tmp_path / f"{name}.{threads}.bin"
            p = subprocess.run([str(exe), str(path), str(out)], capture_output=True, text=True, env=env, timeout=60)
            outcomes[threads] = (p.returncode, p.stdout, out.read_bytes() if p.returncode == 0 else b"")
        ref = outcomes["1"]      # the entry-at-a-time loop
        for threads, got in outcomes.items():
            assert got == ref, (name, threads, got[:2], ref[:2])
        if name in clean:
            assert ref[0] == 0, (name, ref[:2])
            loaders = [oracle.load_mtx] + ([oracle.ref_load_mtx] if oracle.ref_available() else [])
            for load in loaders:
                g = load(str(path))
                k = g["nnz"]
                raw = np.frombuffer(ref[2], dtype=np.uint8)
                assert len(raw) == 12 * k, (name, len(raw), k)
                assert np.array_equal(raw[:4 * k].view(np.int32), g["I"]), name
                assert np.array_equal(raw[4 * k:8 * k].view(np.int32), g["J"]), name
                assert np.array_equal(raw[8 * k:].view(np.uint32), g["V"].view(np.uint32)), name   # bit for bit
                assert ref[1].split()[:3] == [str(g["n_rows"]), str(g["n_cols"]), str(k)], name
    # irregular bodies: whatever the loop decides is the answer (equality above); the failures the reference defines
    # must still be failures
    for name in ("short_file", "zero_index", "value_missing"):
        p = subprocess.run([str(exe), str(tmp_path / f"{name}.mtx"), str(tmp_path / "x.bin")], capture_output=True,
                           text=True, timeout=60)
        assert p.returncode != 0, name


CSR_SRC = r"""
#include <cstdio>
#include <gunrock/io/matrix_market.hxx>
using namespace gunrock;
using namespace memory;
template <typename vec_t>
static void dump(FILE* f, const vec_t& v) { fwrite(v.data(), sizeof(v[0]), v.size(), f); }
int main(int argc, char** argv) {
  io::matrix_market_t<int, int, float> mm;
  auto [props, coo] = mm.load(argv[1]);
  if (argc > 3) coo.row_indices[coo.row_indices.size() / 2] = coo.number_of_rows;   // an index out of range
  format::csr_t<memory_space_t::host, int, int, float> csr;
  csr.from_coo(coo);
  format::csc_t<memory_space_t::host, int, int, float> csc;
  csc.from_csr(csr);
  FILE* f = fopen(argv[2], "wb");
  dump(f, csr.row_offsets); dump(f, csr.column_indices); dump(f, csr.nonzero_values);
  dump(f, csc.column_offsets); dump(f, csc.row_indices); dump(f, csc.nonzero_values);
  fclose(f);
  std::printf("%d %d %d\n", csr.number_of_rows, csr.number_of_columns, csr.number_of_nonzeros);
}
"""


def test_from_coo_and_from_csr_with_host_threads_equal_the_serial_loops(tmp_path):
    """format::csr_t::from_coo / csc_t::from_csr (include/gunrock/formats/formats.hxx; reference formats/csr.hxx:81-140,
    csc.hxx:62-102) run their stable counting sort with all host threads.  Any thread count must give the serial
    loop's arrays bit for bit -- a skewed graph with hub rows, duplicates, self loops and empty rows -- and those
    must be the reference's own from_coo (oracle/_ref) and the C restatement's from_coo / transpose."""
    import numpy as np
    import oracle

    src = tmp_path / "c.cu"
    src.write_text(CSR_SRC)
    exe = tmp_path / "c"
    subprocess.run(["nvcc", "-std=c++17", "-O1", "--extended-lambda", "--expt-relaxed-constexpr",
                    "-gencode", "arch=compute_100a,code=sm_100a", f"-I{ROOT}/include", str(src), "-o", str(exe)],
                   check=True, timeout=600)
    rng = np.random.default_rng(11)
    n_rows, n_cols, nnz = 3000, 2500, 200_000
    r = np.minimum((n_rows * rng.random(nnz) ** 4).astype(np.int64), n_rows - 1)      # hubs at the low ids
    r[r % 11 == 5] = 7                                                                   # empty rows + one fat row
    c = np.minimum((n_cols * rng.random(nnz) ** 2).astype(np.int64), n_cols - 1)
    c[::97] = np.minimum(r[::97], n_cols - 1)                                            # self loops
    r[1::2000], c[1::2000] = r[0::2000][:len(r[1::2000])], c[0::2000][:len(c[1::2000])]  # exact duplicates
    w = (rng.integers(1, 1 << 20, nnz) / 64.0).astype(np.float32)
    mtx = tmp_path / "skew.mtx"
    with open(mtx, "w") as f:
        f.write(f"%%MatrixMarket matrix coordinate real general\n{n_rows} {n_cols} {nnz}\n")
        f.write("\n".join(f"{a + 1} {b + 1} {float(x)!r}" for a, b, x in zip(r, c, w)) + "\n")

    def run(threads, *extra):
        env = dict(os.environ)
        env.pop("GUNROCK_B200_HOST_THREADS", None)
        if threads:
            env["GUNROCK_B200_HOST_THREADS"] = threads
        out = tmp_path / f"csr.{threads}.bin"
        p = subprocess.run([str(exe), str(mtx), str(out), *extra], capture_output=True, text=True, env=env, timeout=120)
        return p.returncode, p.stdout, (out.read_bytes() if p.returncode == 0 else b"")

    serial = run("1")
    assert serial[0] == 0 and serial[1].split() == [str(n_rows), str(n_cols), str(nnz)]
    for threads in ("2", "3", "5", "16", None):
        assert run(threads) == serial, threads
    raw = np.frombuffer(serial[2], dtype=np.uint8)
    cut = np.cumsum([0, 4 * (n_rows + 1), 4 * nnz, 4 * nnz, 4 * (n_cols + 1), 4 * nnz, 4 * nnz])
    ro, ci, vv, co, rj, tv = (raw[a:b] for a, b in zip(cut[:-1], cut[1:]))
    checks = [oracle.csr_from_coo(n_rows, r, c, w)]
    if oracle.ref_available():
        checks.append(oracle.ref_csr_from_coo(n_rows, n_cols, r, c, w))
    for e_ro, e_ci, e_v in checks:
        assert np.array_equal(ro.view(np.int32), e_ro) and np.array_equal(ci.view(np.int32), e_ci)
        assert np.array_equal(vv.view(np.uint32), e_v.view(np.uint32))
    t_ro, t_ci, t_v = oracle.csr_transpose(n_rows, n_cols, ro.view(np.int32), ci.view(np.int32), vv.view(np.float32))
    assert np.array_equal(co.view(np.int32), t_ro) and np.array_equal(rj.view(np.int32), t_ci)
    assert np.array_equal(tv.view(np.uint32), t_v.view(np.uint32))
    # an index outside [0, rows) is an error with any thread count (the reference would write out of bounds)
    for threads in ("1", "4"):
        assert run(threads, "corrupt")[0] != 0


def test_c_abi_host_ingest_and_the_python_module_on_top_of_it(tmp_path):
    """`b2g_mtx_load` / `b2g_csr_from_coo_host` (include/gunrock_b200.h; host-only, no device needed) and
    `gunrock.matrix_market_t().load` / `csr_t().from_coo` that sit on them: same arrays as the C restatement and the
    unmodified reference loader / from_coo, bad files raise instead of exiting the interpreter."""
    import numpy as np
    import pytest
    import oracle
    import gunrock_b200 as gb
    from gunrock_b200 import pygunrock as gunrock

    rng = np.random.default_rng(3)
    n, nnz = 400, 9000
    r, c = rng.integers(1, n + 1, nnz), rng.integers(1, n + 1, nnz)
    w = rng.standard_normal(nnz)
    files = {
        "sym.mtx": "%%MatrixMarket matrix coordinate real symmetric\n% comment\n"
                   f"{n} {n} {nnz}\n" + "\n".join(f"{max(a, b)} {min(a, b)} {float(x)!r}" for a, b, x in zip(r, c, w)) + "\n",
        "pat.mtx": f"%%MatrixMarket matrix coordinate pattern general\n{n} {n} {nnz}\n"
                   + "\n".join(f"{a} {b}" for a, b in zip(r, c)) + "\n",
        "split.mtx": "%%MatrixMarket matrix coordinate real general\n9 9 2\n1 2\n3.5 3\n4 1e2\n",   # irregular body
    }
    for name, text in files.items():
        path = tmp_path / name
        path.write_text(text)
        got = gb.load_mtx(str(path))
        for load in [oracle.load_mtx] + ([oracle.ref_load_mtx] if oracle.ref_available() else []):
            exp = load(str(path))
            assert (got["n_rows"], got["n_cols"], got["nnz"]) == (exp["n_rows"], exp["n_cols"], exp["nnz"]), name
            assert (got["directed"], got["weighted"], got["symmetric"]) == \
                (exp["directed"], exp["weighted"], exp["symmetric"]), name
            assert np.array_equal(got["I"], exp["I"]) and np.array_equal(got["J"], exp["J"]), name
            assert np.array_equal(got["V"].view(np.uint32), exp["V"].view(np.uint32)), name
        # the Python module: loader + from_coo
        props, coo = gunrock.matrix_market_t().load(str(path))
        assert (props.directed, props.weighted, props.symmetric) == (got["directed"], got["weighted"], got["symmetric"])
        csr = gunrock.csr_t().from_coo(coo)
        e_ro, e_ci, e_v = oracle.csr_from_coo(coo.number_of_rows, coo.row_indices, coo.column_indices,
                                              coo.nonzero_values)
        assert np.array_equal(csr.row_offsets, e_ro) and np.array_equal(csr.column_indices, e_ci), name
        assert np.array_equal(csr.nonzero_values.view(np.uint32), e_v.view(np.uint32)), name
        assert csr.number_of_nonzeros == got["nnz"] and csr.row_offsets.dtype == np.int32
    # from_coo without values, an empty COO, an out-of-range row
    ro, ci, vals = gb.csr_from_coo_host(5, [4, 0, 4, 2], [1, 2, 0, 2])
    assert ro.tolist() == [0, 1, 1, 2, 2, 4] and ci.tolist() == [2, 2, 1, 0] and vals is None
    ro, ci, vals = gb.csr_from_coo_host(3, [], [], np.zeros(0, np.float32))
    assert ro.tolist() == [0, 0, 0, 0] and ci.size == 0 and vals.size == 0
    with pytest.raises(gb.GunrockB200Error, match="out of range"):
        gb.csr_from_coo_host(3, [0, 3], [0, 0])
    # bad files: an error with the reference's message, not exit(1)
    for text, message in (("%%MatrixMarket matrix array real general\n2 2\n1\n2\n3\n4\n", "not a sparse matrix"),
                          ("hello\n", "banner"),
                          ("%%MatrixMarket matrix coordinate complex general\n2 2 1\n1 1 1 0\n", "format type"),
                          ("%%MatrixMarket matrix coordinate pattern general\n2 2 2\n1 1\n", "Could not read edge"),
                          ("%%MatrixMarket matrix coordinate pattern general\n2 2 1\n0 1\n", "zero-indexed")):
        bad = tmp_path / "bad.mtx"
        bad.write_text(text)
        with pytest.raises(gb.GunrockB200Error, match=message):
            gb.load_mtx(str(bad))
    with pytest.raises(gb.GunrockB200Error, match="could not be opened"):
        gb.load_mtx(str(tmp_path / "missing.mtx"))
