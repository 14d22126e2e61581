// oracle/ref_gpu_driver.cu -- TEST INFRASTRUCTURE ONLY (never shipped, never on the product path).
//
// The UNMODIFIED reference GPU implementation (gunrock::{bfs,sssp,pr}::run and the operators under
// /root/reference/include) compiled for sm_100a from where its sources lie, behind a small command
// line so that bench.py can time "the reference's own kernels on this very GPU" next to ours
// (optional second baseline row).  Built by `make -C oracle ref_gpu` into
// oracle/_ref/gunrock_ref_gpu with -include ref_gpu_fix.h and -DSM_TARGET=90 (F3).
//
//   gunrock_ref_gpu <bfs|sssp|pr> <graph.csr> <source> <runs> <load_balance> [validate] [dump=<file>]
//
// `dump=<file>` writes the result array of the last run (int32 depths / fp32 distances / fp32 ranks) as raw
// bytes: tests/test_zz_gpu_widening.py compares our results with what the reference's OWN GPU kernels
// produce on the same GPU (the only PageRank output of the reference there is).
//
// graph.csr is the reference's own binary layout (formats/csr.hxx:142-228).  Prints one JSON line:
// per-run milliseconds as returned by run() (the enactor's own timer, enactor.hxx:266-288) and, with
// `validate`, the number of mismatches against the reference CPU validators.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include <gunrock/algorithms/bfs.hxx>
#include <gunrock/algorithms/pr.hxx>
#include <gunrock/algorithms/sssp.hxx>

#include "bfs_cpu.hxx"
#include "sssp_cpu.hxx"

using namespace gunrock;
using namespace memory;

using vertex_t = int;
using edge_t = int;
using weight_t = float;

int main(int argc, char** argv) {
  if (argc < 6) {
    std::fprintf(stderr, "usage: %s <bfs|sssp|pr> <graph.csr> <source> <runs> <load_balance> [validate]\n", argv[0]);
    return 2;
  }
  const std::string alg = argv[1], file = argv[2], lb_name = argv[5];
  vertex_t source = std::atoi(argv[3]);  // the validators take it by non-const reference
  const int runs = std::atoi(argv[4]);
  bool validate = false;
  std::string dump_path;
  for (int i = 6; i < argc; ++i) {
    std::string a = argv[i];
    if (a == "validate")
      validate = true;
    else if (a.rfind("dump=", 0) == 0)
      dump_path = a.substr(5);
  }
  auto dump = [&](const void* host_ptr, size_t bytes) {
    if (dump_path.empty())
      return;
    FILE* f = std::fopen(dump_path.c_str(), "wb");
    if (!f || std::fwrite(host_ptr, 1, bytes, f) != bytes) {
      std::fprintf(stderr, "cannot write %s\n", dump_path.c_str());
      std::exit(3);
    }
    std::fclose(f);
  };

  using csr_t = format::csr_t<memory_space_t::device, vertex_t, edge_t, weight_t>;
  // The file is in the layout of the reference's csr_t::write_binary (formats/csr.hxx:193-228) but is read
  // here, not by csr_t::read_binary: that function wraps every fread in `throw_if_exception(fread(...) != 0)`
  // (formats/csr.hxx:146-151), i.e. it throws exactly when a read SUCCEEDS.
  csr_t csr;
  {
    FILE* f = std::fopen(file.c_str(), "rb");
    int header[3];
    if (!f || std::fread(header, sizeof(int), 3, f) != 3) {
      std::fprintf(stderr, "cannot read %s\n", file.c_str());
      return 2;
    }
    thrust::host_vector<edge_t> ro((size_t)header[0] + 1);
    thrust::host_vector<vertex_t> ci((size_t)header[2]);
    thrust::host_vector<weight_t> vals((size_t)header[2]);
    bool ok = std::fread(ro.data(), sizeof(edge_t), ro.size(), f) == ro.size() &&
              std::fread(ci.data(), sizeof(vertex_t), ci.size(), f) == ci.size() &&
              std::fread(vals.data(), sizeof(weight_t), vals.size(), f) == vals.size();
    std::fclose(f);
    if (!ok) {
      std::fprintf(stderr, "truncated %s\n", file.c_str());
      return 2;
    }
    csr.number_of_rows = header[0];
    csr.number_of_columns = header[1];
    csr.number_of_nonzeros = header[2];
    csr.row_offsets = ro;
    csr.column_indices = ci;
    csr.nonzero_values = vals;
  }
  graph::graph_properties_t properties;
  auto G = graph::build<memory_space_t::device>(properties, csr);
  const vertex_t V = G.get_number_of_vertices();

  gunrock::options_t options;
  if (lb_name == "merge_path")
    options.advance_load_balance = operators::load_balance_t::merge_path;
  else if (lb_name == "thread_mapped")
    options.advance_load_balance = operators::load_balance_t::thread_mapped;
  else
    options.advance_load_balance = operators::load_balance_t::block_mapped;

  auto context = std::make_shared<gcuda::multi_context_t>(0);
  std::vector<float> ms;
  long long errors = -1;
  if (alg == "bfs") {
    thrust::device_vector<vertex_t> distances(V), predecessors(V);
    for (int r = 0; r < runs; ++r) {
      bfs::param_t<vertex_t> param(source, options);
      bfs::result_t<vertex_t> result(distances.data().get(), predecessors.data().get());
      ms.push_back(bfs::run(G, param, result, context));
    }
    {
      thrust::host_vector<vertex_t> h(distances);
      dump(h.data(), sizeof(vertex_t) * (size_t)V);
    }
    if (validate) {
      thrust::host_vector<vertex_t> h(distances), exp(V), pred(V);
      bfs_cpu::run<csr_t, vertex_t, edge_t>(csr, source, exp.data(), pred.data());
      errors = 0;
      for (vertex_t v = 0; v < V; ++v)
        errors += h[v] != exp[v];
    }
  } else if (alg == "sssp") {
    thrust::device_vector<weight_t> distances(V);
    thrust::device_vector<vertex_t> predecessors(V);
    for (int r = 0; r < runs; ++r) {
      sssp::param_t<vertex_t> param(source, options);
      sssp::result_t<vertex_t, weight_t> result(distances.data().get(), predecessors.data().get(), V);
      ms.push_back(sssp::run(G, param, result, context));
    }
    {
      thrust::host_vector<weight_t> h(distances);
      dump(h.data(), sizeof(weight_t) * (size_t)V);
    }
    if (validate) {
      thrust::host_vector<weight_t> h(distances), exp(V);
      thrust::host_vector<vertex_t> pred(V);
      sssp_cpu::run<csr_t, vertex_t, edge_t, weight_t>(csr, source, exp.data(), pred.data());
      errors = 0;
      for (vertex_t v = 0; v < V; ++v)
        errors += h[v] != exp[v];
    }
  } else if (alg == "pr") {
    thrust::device_vector<weight_t> p(V);
    for (int r = 0; r < runs; ++r) {
      pr::param_t<weight_t> param(0.85f, 1e-6f, options);
      pr::result_t<weight_t> result(p.data().get());
      ms.push_back(pr::run(G, param, result, context));
    }
    thrust::host_vector<weight_t> h(p);
    dump(h.data(), sizeof(weight_t) * (size_t)V);
  } else {
    std::fprintf(stderr, "unknown algorithm %s\n", alg.c_str());
    return 2;
  }
  context->get_context(0)->synchronize();
  std::printf("{\"impl\": \"reference_gpu\", \"algorithm\": \"%s\", \"load_balance\": \"%s\", \"vertices\": %d, "
              "\"edges\": %d, \"source\": %d, \"errors\": %lld, \"ms\": [",
              alg.c_str(), lb_name.c_str(), (int)V, (int)G.get_number_of_edges(), (int)source, errors);
  for (size_t i = 0; i < ms.size(); ++i)
    std::printf("%s%.4f", i ? ", " : "", ms[i]);
  std::printf("]}\n");
  return 0;
}
