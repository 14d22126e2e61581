/**
 * @file performance.hxx
 * @brief `util::stats::export_performance_stats(13 args)` -- the JSON run report of the examples
 * (include/gunrock/util/performance.hxx:82-285).  Same keys (engine, primitive, graph_type,
 * num_edges, num_vertices, srcs, tags, process_times, avg/min/max/stddev_process_time, time,
 * nodes_visited/edges_visited/search_depths and mteps = edges_visited / ms / 1000 (:225-231), gpuinfo,
 * command_line); written with a small built-in JSON emitter (nlohmann/json is a network-fetched
 * dependency of the reference).  Unlike the reference the metrics need no special build: the
 * kernels always account visited edges (framework/benchmark.hxx).
 */
#pragma once

#include <cmath>
#include <ctime>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <numeric>
#include <sstream>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include <gunrock/framework/benchmark.hxx>
#include <gunrock/util/filepath.hxx>

namespace gunrock {
namespace util {
namespace stats {

namespace detail {

inline std::string quote(const std::string& s) {
  std::string out = "\"";
  for (char c : s) {
    if (c == '"' || c == '\\')
      out += '\\';
    if (c == '\n')
      out += "\\n";
    else
      out += c;
  }
  return out + "\"";
}

template <typename T>
std::string array(const std::vector<T>& v) {
  std::ostringstream os;
  os << std::setprecision(9) << "[";
  for (std::size_t i = 0; i < v.size(); ++i)
    os << (i ? ", " : "") << v[i];
  os << "]";
  return os.str();
}
inline std::string array(const std::vector<std::string>& v) {
  std::string out = "[";
  for (std::size_t i = 0; i < v.size(); ++i)
    out += (i ? ", " : "") + quote(v[i]);
  return out + "]";
}

struct writer_t {
  std::ostringstream os;
  bool first = true;
  writer_t() { os << std::setprecision(9) << "{\n"; }
  void raw(const std::string& key, const std::string& value) {
    os << (first ? "" : ",\n") << "    " << quote(key) << ": " << value;
    first = false;
  }
  template <typename T>
  void num(const std::string& key, T v) {
    std::ostringstream t;
    t << std::setprecision(9) << v;
    raw(key, t.str());
  }
  void str(const std::string& key, const std::string& v) { raw(key, quote(v)); }
  std::string finish() {
    os << "\n}\n";
    return os.str();
  }
};

template <typename T>
void min_max_avg(const std::vector<T>& v, double& mn, double& mx, double& avg, double& sd) {
  mn = mx = avg = sd = 0;
  if (v.empty())
    return;
  mn = mx = static_cast<double>(v[0]);
  double sum = 0;
  for (auto& x : v) {
    double d = static_cast<double>(x);
    mn = d < mn ? d : mn;
    mx = d > mx ? d : mx;
    sum += d;
  }
  avg = sum / v.size();
  double acc = 0;
  for (auto& x : v)
    acc += (static_cast<double>(x) - avg) * (static_cast<double>(x) - avg);
  sd = v.size() > 1 ? std::sqrt(acc / (v.size() - 1)) : 0.0;
}

}  // namespace detail

inline void export_performance_stats(std::vector<benchmark::host_benchmark_t>& benchmark_metrics,
                                     size_t edges,
                                     size_t vertices,
                                     std::vector<float>& run_times,
                                     std::string primitive,
                                     std::string filename,
                                     std::string graph_type,
                                     std::string json_dir,
                                     std::string json_file,
                                     std::vector<int>& sources,
                                     std::vector<std::string>& tags,
                                     int argc,
                                     char** argv) {
  detail::writer_t j;
  double mn, mx, avg, sd;
  detail::min_max_avg(run_times, mn, mx, avg, sd);

  std::time_t now = std::time(nullptr);
  std::string time_s = std::ctime(&now);
  if (!time_s.empty() && time_s.back() == '\n')
    time_s.pop_back();

  std::string command_line;
  for (int i = 0; i < argc; ++i)
    command_line += std::string(i ? " " : "") + argv[i];

  j.str("engine", "Essentials (B200-native)");
  j.str("primitive", primitive);
  j.str("graph_type", graph_type);
  j.num("num_edges", edges);
  j.num("num_vertices", vertices);
  j.raw("srcs", detail::array(sources));
  j.raw("tags", detail::array(tags));
  j.raw("process_times", detail::array(run_times));
  j.str("graph_file", util::extract_filename(filename));
  j.num("avg_process_time", avg);
  j.num("stddev_process_time", sd);
  j.num("min_process_time", mn);
  j.num("max_process_time", mx);
  j.str("time", time_s);
  j.str("command_line", command_line);

  std::vector<unsigned long long> nodes, edgs;
  std::vector<int> depths;
  std::vector<double> mteps;
  for (std::size_t i = 0; i < benchmark_metrics.size(); ++i) {
    nodes.push_back(benchmark_metrics[i].vertices_visited);
    edgs.push_back(benchmark_metrics[i].edges_visited);
    depths.push_back(benchmark_metrics[i].search_depth);
    double ms = i < run_times.size() ? run_times[i] : 0.0;
    mteps.push_back(ms > 0 ? benchmark_metrics[i].edges_visited / ms / 1000.0 : 0.0);
  }
  j.raw("nodes_visited", detail::array(nodes));
  j.raw("edges_visited", detail::array(edgs));
  j.raw("search_depths", detail::array(depths));
  detail::min_max_avg(depths, mn, mx, avg, sd);
  j.num("avg_search_depth", avg);
  j.num("min_search_depth", mn);
  j.num("max_search_depth", mx);
  j.raw("mteps", detail::array(mteps));
  detail::min_max_avg(mteps, mn, mx, avg, sd);
  j.num("avg_mteps", avg);
  j.num("min_mteps", mn);
  j.num("max_mteps", mx);

  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&prop, dev) == cudaSuccess) {
    int driver = 0, runtime = 0;
    cudaDriverGetVersion(&driver);
    cudaRuntimeGetVersion(&runtime);
    std::ostringstream g;
    g << "{" << detail::quote("name") << ": " << detail::quote(prop.name) << ", "
      << detail::quote("total_global_mem") << ": " << prop.totalGlobalMem << ", "
      << detail::quote("major") << ": " << prop.major << ", " << detail::quote("minor") << ": "
      << prop.minor << ", " << detail::quote("multi_processor_count") << ": "
      << prop.multiProcessorCount << ", " << detail::quote("driver_api_version") << ": " << driver
      << ", " << detail::quote("runtime_api_version") << ": " << runtime << "}";
    j.raw("gpuinfo", g.str());
  }

  std::string path;
  if (json_file == "") {
    std::string stamp = time_s;
    for (auto& c : stamp)
      if (c == ' ')
        c = '_';
    stamp.erase(std::remove(stamp.begin(), stamp.end(), ':'), stamp.end());
    path = json_dir + "/" + primitive + "_" + util::extract_filename(filename) + "_" + stamp + ".json";
  } else {
    path = json_dir + "/" + json_file;
  }
  std::ofstream out(path);
  out << j.finish();
}

}  // namespace stats
}  // namespace util
}  // namespace gunrock
