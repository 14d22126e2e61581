// Gate 2 for the drop-in boundary: the REFERENCE's own algorithm headers
// (/root/reference/include/gunrock/algorithms/{bfs,sssp,pr}.hxx, included by absolute path and
// unmodified) compiled against THIS repository's framework/operator headers, and run:
//   ref_algorithms_driver <graph.mtx> <source>
// prints the BFS/SSSP error counts against the reference CPU validators and a PageRank digest.
// This proves user-written algorithms (lambdas over advance / filter / parallel_for, enactor_t,
// problem_t, frontier_t) keep working on the B200 operators.  Built only where /root/reference
// exists (examples/build_reference_examples.sh).
#include <gunrock/algorithms/algorithms.hxx>  // ours (framework, operators, formats, io)

#include REF_BFS_HXX   // e.g. "/root/reference/include/gunrock/algorithms/bfs.hxx"
#include REF_SSSP_HXX
#include REF_PR_HXX

#include "bfs_cpu.hxx"
#include "sssp_cpu.hxx"

using namespace gunrock;
using namespace memory;

int main(int argc, char** argv) {
  if (argc < 3) {
    std::cerr << "usage: " << argv[0] << " graph.mtx source [load_balance]" << std::endl;
    return 2;
  }
  using vertex_t = int;
  using edge_t = int;
  using weight_t = float;
  using csr_t = format::csr_t<memory_space_t::device, vertex_t, edge_t, weight_t>;

  io::matrix_market_t<vertex_t, edge_t, weight_t> mm;
  auto [properties, coo] = mm.load(argv[1]);
  csr_t csr;
  csr.from_coo(coo);
  auto G = graph::build<memory_space_t::device>(properties, csr);
  vertex_t source = std::atoi(argv[2]);
  auto context = std::make_shared<gcuda::multi_context_t>(0);
  size_t n = G.get_number_of_vertices();

  gunrock::options_t options;
  if (argc > 3) {
    std::string lb = argv[3];
    options.advance_load_balance = lb == "merge_path"      ? operators::load_balance_t::merge_path
                                   : lb == "thread_mapped" ? operators::load_balance_t::thread_mapped
                                                           : operators::load_balance_t::block_mapped;
  }
  options.enable_filter = true;

  {
    thrust::device_vector<vertex_t> distances(n), predecessors(n);
    gunrock::bfs::param_t<vertex_t> param(source, options);
    gunrock::bfs::result_t<vertex_t> result(distances.data().get(), predecessors.data().get());
    float ms = gunrock::bfs::run(G, param, result, context);
    thrust::host_vector<vertex_t> h(n), hp(n);
    bfs_cpu::run<csr_t, vertex_t, edge_t>(csr, source, h.data(), hp.data());
    int errors = util::compare(distances.data().get(), h.data(), n);
    std::cout << "ref-bfs.hxx on B200 operators: " << ms << " ms, errors : " << errors << std::endl;
  }
  {
    thrust::device_vector<weight_t> distances(n);
    thrust::device_vector<vertex_t> predecessors(n);
    options.enable_uniquify = true;
    gunrock::sssp::param_t<vertex_t> param(source, options);
    gunrock::sssp::result_t<vertex_t, weight_t> result(distances.data().get(),
                                                       predecessors.data().get(), n);
    float ms = gunrock::sssp::run(G, param, result, context);
    thrust::host_vector<weight_t> h(n);
    thrust::host_vector<vertex_t> hp(n);
    sssp_cpu::run<csr_t, vertex_t, edge_t, weight_t>(csr, source, h.data(), hp.data());
    int errors = util::compare(distances.data().get(), h.data(), n);
    std::cout << "ref-sssp.hxx on B200 operators: " << ms << " ms, errors : " << errors << std::endl;
  }
  {
    thrust::device_vector<weight_t> p(n);
    gunrock::pr::param_t<weight_t> param(0.85f, 1e-6f, options);
    gunrock::pr::result_t<weight_t> result(p.data().get());
    float ms = gunrock::pr::run(G, param, result, context);
    thrust::host_vector<weight_t> h(p);
    double sum = 0;
    for (auto x : h)
      sum += x;
    std::cout << "ref-pr.hxx on B200 operators: " << ms << " ms, sum : " << sum << std::endl;
    std::cout << "ranks :";
    for (size_t i = 0; i < n && i < 64; ++i)
      std::cout << " " << h[i];
    std::cout << std::endl;
  }
  return 0;
}
