"""Mint the golden vectors under tests/golden/ from the UNMODIFIED reference (oracle/_ref).

Run in the authoring container (needs /root/reference and `make -C oracle ref`):
    python tests/golden/make_golden.py
Every expected array below is produced by the reference's own code compiled from where it lies:
    io::matrix_market_t::load, io::smtx_t::load, format::csr_t::from_coo, bfs_cpu::run, sssp_cpu::run
(oracle/ref_driver.cu).  The RMAT inputs are defined by the oracle's counter-based generator; only
its parameters and a sha256 of the resulting CSR are stored, the expected distances come from
the reference validators run on that CSR.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import oracle  # noqa: E402

REF = "/root/reference"


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def f32_hex(a):
    """fp32 arrays are stored as uint32 bit patterns: bit-exact and JSON-safe (FLT_MAX included)."""
    return np.ascontiguousarray(a, np.float32).view(np.uint32).tolist()


def main():
    out = {}
    # 1. chesapeake.mtx through the reference loader + from_coo + validators
    path = os.path.join(REF, "datasets/chesapeake/chesapeake.mtx")
    m = oracle.ref_load_mtx(path)
    ro, ci, v = oracle.ref_csr_from_coo(m["n_rows"], m["n_cols"], m["I"], m["J"], m["V"])
    g = oracle.RefGraph(ro, ci, v)
    ches = dict(n_rows=m["n_rows"], nnz=m["nnz"], directed=m["directed"], weighted=m["weighted"],
                symmetric=m["symmetric"], row_offsets=ro.tolist(), column_indices=ci.tolist(),
                values_bits=f32_hex(v), coo_I=m["I"].tolist(), coo_J=m["J"].tolist(), bfs={}, sssp={})
    for src in (0, 7, 38):
        ches["bfs"][str(src)] = g.bfs(src)[0].tolist()
        ches["sssp"][str(src)] = f32_hex(g.sssp(src)[0])
    out["chesapeake"] = ches

    # 2. the reference's pytest fixture graph (python/tests/conftest.py:17-34)
    I = np.array([0, 0, 1, 2, 3], np.int32)
    J = np.array([1, 2, 3, 3, 4], np.int32)
    V = np.array([1.0, 2.0, 1.5, 1.0, 2.5], np.float32)
    ro, ci, v = oracle.ref_csr_from_coo(5, 5, I, J, V)
    g = oracle.RefGraph(ro, ci, v)
    out["pytest_dag"] = dict(I=I.tolist(), J=J.tolist(), V_bits=f32_hex(V), row_offsets=ro.tolist(),
                             column_indices=ci.tolist(), bfs={"0": g.bfs(0)[0].tolist()},
                             sssp={"0": f32_hex(g.sssp(0)[0])})

    # 3. the 4x4 sample CSR (include/gunrock/io/sample.hxx:57-92)
    ro = np.array([0, 0, 2, 3, 4], np.int32)
    ci = np.array([0, 1, 2, 1], np.int32)
    v = np.array([5, 8, 3, 6], np.float32)
    g = oracle.RefGraph(ro, ci, v)
    out["sample4"] = dict(row_offsets=ro.tolist(), column_indices=ci.tolist(), values_bits=f32_hex(v),
                          bfs={str(s): g.bfs(s)[0].tolist() for s in range(4)},
                          sssp={str(s): f32_hex(g.sssp(s)[0]) for s in range(4)})

    # 4. seeded RMAT graphs (generator = workload definition, expected = reference validators)
    for name, scale, ef, seed, wseed, nonint in (("rmat10", 10, 16, 0x5EED10, 11, False),
                                                 ("rmat12w", 12, 8, 0x5EED12, 12, True)):
        ro, ci = oracle.rmat_csr(scale, ef, seed)
        w = oracle.edge_weights(wseed, ro, ci, nonint)
        g = oracle.RefGraph(ro, ci, w)
        deg = np.diff(ro)
        srcs = [int(deg.argmax()), 1, int(np.flatnonzero(deg > 0)[-1])]
        out[name] = dict(scale=scale, edge_factor=ef, seed=seed, weight_seed=wseed,
                         non_integer=nonint, nnz=int(len(ci)), csr_sha256=sha(ro, ci),
                         weights_sha256=sha(w), sources=srcs,
                         bfs={str(s): g.bfs(s)[0].tolist() for s in srcs},
                         sssp={str(s): f32_hex(g.sssp(s)[0]) for s in srcs})
    # 5. the reference's second vendored dataset: bips98_606.mtx (real general, 7135 vertices, 34738
    #    entries, DIRECTED, explicit diagonal, negative values).  Loader + from_coo + bfs_cpu as they are;
    #    for SSSP the weights are |value| (the validator, like the GPU algorithm, needs weights >= 0).
    #    Stored as a compressed .npz next to this file (the arrays would triple golden.json).
    path = os.path.join(REF, "datasets/bips98_606/bips98_606.mtx")
    m = oracle.ref_load_mtx(path)
    ro, ci, v = oracle.ref_csr_from_coo(m["n_rows"], m["n_cols"], m["I"], m["J"], m["V"])
    wabs = np.abs(v).astype(np.float32)
    g = oracle.RefGraph(ro, ci, wabs)
    deg = np.diff(ro)
    srcs = [0, int(deg.argmax()), int(m["n_rows"] - 1)]
    arrays = dict(row_offsets=ro, column_indices=ci, values_bits=v.view(np.uint32),
                  props=np.array([m["directed"], m["weighted"], m["symmetric"]], np.int32),
                  sources=np.array(srcs, np.int32))
    for s in srcs:
        arrays[f"bfs_{s}"] = g.bfs(s)[0]
        arrays[f"sssp_abs_bits_{s}"] = np.ascontiguousarray(g.sssp(s)[0], np.float32).view(np.uint32)
    np.savez_compressed(os.path.join(HERE, "bips98_606.npz"), **arrays)
    print("wrote bips98_606.npz", os.path.getsize(os.path.join(HERE, "bips98_606.npz")), "bytes; sources", srcs)

    # 6. the reference's `.smtx` dataset through its own reader (io/smtx.hxx:56-169); its unit test asserts the
    #    sizes 96 x 96 with 4608 entries (unittests/io/smtx.cuh:28-30).  Only the structure is stored: the values
    #    the reference attaches are random draws.
    path = os.path.join(REF, "datasets/layers.0.blocks.0.attn.proj_swin_tiny_unstructured_50.smtx")
    g = oracle.ref_load_smtx(path)
    assert (g["n_rows"], g["n_cols"], g["nnz"]) == (96, 96, 4608)
    np.savez_compressed(os.path.join(HERE, "swin_tiny_attn_proj_smtx.npz"),
                        shape=np.array([g["n_rows"], g["n_cols"], g["nnz"]], np.int32),
                        row_offsets=g["row_offsets"], column_indices=g["column_indices"])
    print("wrote swin_tiny_attn_proj_smtx.npz")

    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", os.path.join(HERE, "golden.json"), os.path.getsize(os.path.join(HERE, "golden.json")), "bytes")


if __name__ == "__main__":
    main()
