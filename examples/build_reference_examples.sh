#!/bin/bash
# Gates for the drop-in boundary (need /root/reference; binaries land in examples/bin/, git-ignored,
# and travel to the GPU box with the snapshot):
#  1. the reference's UNCHANGED example translation units examples/algorithms/{bfs,sssp,pr}/*.cu
#     compiled against THIS repository's headers            -> bin/{bfs,sssp,pr}
#  2. the same with -DGUNROCK_B200_OPERATOR_PATH (our algorithm headers on the generic operators
#     instead of the fused enactors)                        -> bin/{bfs,sssp,pr}_ops
#  3. the reference's own algorithm headers (algorithms/{bfs,sssp,pr}.hxx) on our framework /
#     operator headers                                      -> bin/ref_algorithms
#  0. (no reference needed) examples/api_selftest.cu -> bin/api_selftest,
#     examples/dense_frontier_selftest.cu (bitmap / boolmap frontier views) -> bin/dense_frontier_selftest,
#     examples/cache_selftest.cu (per-graph caches of the fused enactors across graphs / contexts) -> bin/cache_selftest,
#     examples/multi_context_selftest.cu (bfs::run over a multi-device gcuda::multi_context_t) -> bin/multi_context_selftest,
#     examples/reference_layout_selftest.cu (opt-in: advance output in the reference's edge-rank layout) -> bin/reference_layout_selftest
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF=${REF:-/root/reference}
OUT="$ROOT/examples/bin"
mkdir -p "$OUT"
FLAGS="-std=c++17 -O3 -lineinfo --extended-lambda --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a -I$ROOT/include -diag-suppress 20050"
pids=()
nvcc $FLAGS -o "$OUT/api_selftest" "$ROOT/examples/api_selftest.cu" & pids+=($!)
nvcc $FLAGS -o "$OUT/dense_frontier_selftest" "$ROOT/examples/dense_frontier_selftest.cu" & pids+=($!)
nvcc $FLAGS -o "$OUT/cache_selftest" "$ROOT/examples/cache_selftest.cu" & pids+=($!)
nvcc $FLAGS -o "$OUT/multi_context_selftest" "$ROOT/examples/multi_context_selftest.cu" & pids+=($!)
nvcc $FLAGS -o "$OUT/reference_layout_selftest" "$ROOT/examples/reference_layout_selftest.cu" & pids+=($!)
if [ ! -d "$REF" ]; then
  rc=0; for p in "${pids[@]}"; do wait $p || rc=1; done; exit $rc
fi
for alg in bfs sssp pr; do
  nvcc $FLAGS -I"$REF/examples/algorithms/$alg" -o "$OUT/$alg" "$REF/examples/algorithms/$alg/$alg.cu" & pids+=($!)
  nvcc $FLAGS -DGUNROCK_B200_OPERATOR_PATH -I"$REF/examples/algorithms/$alg" -o "$OUT/${alg}_ops" "$REF/examples/algorithms/$alg/$alg.cu" & pids+=($!)
done
A="$REF/include/gunrock/algorithms"
nvcc $FLAGS -I"$REF/examples/algorithms/bfs" -I"$REF/examples/algorithms/sssp" \
  -DREF_BFS_HXX="\"$A/bfs.hxx\"" -DREF_SSSP_HXX="\"$A/sssp.hxx\"" -DREF_PR_HXX="\"$A/pr.hxx\"" \
  -o "$OUT/ref_algorithms" "$ROOT/examples/ref_algorithms_driver.cu" & pids+=($!)
#  4. widening: the reference's OTHER algorithms -- their own algorithm headers and
#     unchanged example TUs -- on our framework/operator headers -> bin/ext_<alg>
#     (mst / tc need <cxxopts.hpp>: served by include/cxxopts.hpp; tc runs with uint32_t vertex / edge ids;
#      spgemm uses the csr + csc multi-view graph and the view-tagged accessors of graph/graph.hxx)
SHIM="$OUT/shim/gunrock/algorithms"
mkdir -p "$SHIM"
for alg in bc color geo hits kcore ppr spmv mst tc spgemm; do
  echo "#include \"$REF/include/gunrock/algorithms/$alg.hxx\"" > "$SHIM/$alg.hxx"
  nvcc $FLAGS -I"$OUT/shim" -I"$REF/examples/algorithms/$alg" -o "$OUT/ext_$alg" "$REF/examples/algorithms/$alg/$alg.cu" & pids+=($!)
done
#  5. the reference's tools, unchanged: tools/csr_binary.cu (.mtx -> binary .csr through our loader, from_coo and
#     write_binary) -> bin/tool_csr_binary, tools/cmd.cu (the cxxopts surface of include/cxxopts.hpp) -> bin/tool_cmd
nvcc $FLAGS -o "$OUT/tool_csr_binary" "$REF/examples/tools/csr_binary.cu" & pids+=($!)
nvcc $FLAGS -o "$OUT/tool_cmd" "$REF/examples/tools/cmd.cu" & pids+=($!)
#  6. the reference's OWN unit tests (unittests/unittests.hxx and the 15 headers it includes), unchanged, against
#     our include tree; tests/gtest_shim/gtest/gtest.h stands in for googletest (not in the image); tc.hxx is the
#     reference's, through the shim directory of section 4 -> bin/ref_unittests (--gtest_filter, --gtest_list_tests)
nvcc $FLAGS -I"$OUT/shim" -I"$REF/unittests" -I"$ROOT/tests/gtest_shim" -o "$OUT/ref_unittests" "$ROOT/examples/ref_unittests_main.cu" & pids+=($!)
rc=0
for p in "${pids[@]}"; do wait $p || rc=1; done
ls -la "$OUT"
exit $rc
