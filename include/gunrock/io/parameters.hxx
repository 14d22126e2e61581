/**
 * @file parameters.hxx
 * @brief `io::cli::parameters_t` and the source / tag string parsers
 * (include/gunrock/io/parameters.hxx:16-236).  Same flags, defaults and exit behaviour; the option
 * parser is a small built-in one (cxxopts is a network-fetched dependency of the reference).
 *   -m/--market FILE  -s/--src LIST  -n/--num_runs N  --validate  --export_metrics
 *   -d/--json_dir DIR -f/--json_file FILE -t/--tag LIST
 *   --advance_load_balance {thread_mapped,block_mapped,merge_path,merge_path_v2,...}
 *   --filter_algorithm {remove,predicated,compact,bypass} --enable_filter
 *   --enable_uniquify --uniquify_algorithm {unique,unique_copy} --best_effort_uniquify
 *   --uniquify_percent P
 * B200 addition: --advance_direction {forward,backward,optimized}.
 */
#pragma once

#include <algorithm>
#include <cstdlib>
#include <iostream>
#include <map>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include <gunrock/algorithms/algorithms.hxx>
#include <gunrock/framework/operators/configs.hxx>
#include <gunrock/util/filepath.hxx>

namespace gunrock {
namespace io {
namespace cli {

inline operators::load_balance_t parse_load_balance(std::string str) {
  std::transform(str.begin(), str.end(), str.begin(), ::tolower);
  if (str == "thread_mapped")
    return operators::load_balance_t::thread_mapped;
  if (str == "warp_mapped")
    return operators::load_balance_t::warp_mapped;
  if (str == "block_mapped")
    return operators::load_balance_t::block_mapped;
  if (str == "bucketing")
    return operators::load_balance_t::bucketing;
  if (str == "merge_path")
    return operators::load_balance_t::merge_path;
  if (str == "merge_path_v2")
    return operators::load_balance_t::merge_path_v2;
  if (str == "work_stealing")
    return operators::load_balance_t::work_stealing;
  std::cerr << "Warning: Unknown load_balance type '" << str << "', using block_mapped"
            << std::endl;
  return operators::load_balance_t::block_mapped;
}

inline operators::filter_algorithm_t parse_filter_algorithm(std::string str) {
  std::transform(str.begin(), str.end(), str.begin(), ::tolower);
  if (str == "remove")
    return operators::filter_algorithm_t::remove;
  if (str == "predicated")
    return operators::filter_algorithm_t::predicated;
  if (str == "compact")
    return operators::filter_algorithm_t::compact;
  if (str == "bypass")
    return operators::filter_algorithm_t::bypass;
  std::cerr << "Warning: Unknown filter algorithm '" << str << "', using predicated" << std::endl;
  return operators::filter_algorithm_t::predicated;
}

inline operators::uniquify_algorithm_t parse_uniquify_algorithm(std::string str) {
  std::transform(str.begin(), str.end(), str.begin(), ::tolower);
  if (str == "unique_copy")
    return operators::uniquify_algorithm_t::unique_copy;
  if (str != "unique")
    std::cerr << "Warning: Unknown uniquify algorithm '" << str << "', using unique" << std::endl;
  return operators::uniquify_algorithm_t::unique;
}

inline operators::advance_direction_t parse_advance_direction(std::string str) {
  std::transform(str.begin(), str.end(), str.begin(), ::tolower);
  if (str == "backward" || str == "pull")
    return operators::advance_direction_t::backward;
  if (str == "optimized" || str == "optimised")
    return operators::advance_direction_t::optimized;
  return operators::advance_direction_t::forward;
}

struct parameters_t {
  std::string filename;
  std::string source_string = "";
  std::string json_dir = ".";
  std::string json_file = "";
  std::string tag_string = "";
  int num_runs = 1;
  bool export_metrics = false;
  bool validate = false;
  bool binary = false;

  operators::load_balance_t advance_load_balance = operators::load_balance_t::block_mapped;
  operators::filter_algorithm_t filter_algorithm = operators::filter_algorithm_t::predicated;
  bool enable_filter = false;
  bool enable_uniquify = false;
  operators::uniquify_algorithm_t uniquify_algorithm = operators::uniquify_algorithm_t::unique;
  bool best_effort_uniquify = true;
  float uniquify_percent = 100.0f;
  operators::advance_direction_t advance_direction = operators::advance_direction_t::forward;

  parameters_t(int argc, char** argv, std::string algorithm) {
    const bool has_sources = algorithm == "Betweenness Centrality" ||
                             algorithm == "Breadth First Search" ||
                             algorithm == "Single Source Shortest Path";
    const bool has_validate =
        algorithm == "Breadth First Search" || algorithm == "Single Source Shortest Path";
    program = argc > 0 ? argv[0] : "gunrock";
    description = algorithm + " example";

    // long name -> takes a value?
    std::map<std::string, bool> known = {
        {"help", false}, {"export_metrics", false}, {"market", true}, {"json_dir", true},
        {"json_file", true}, {"tag", true}, {"advance_load_balance", true},
        {"filter_algorithm", true}, {"enable_filter", false}, {"enable_uniquify", false},
        {"uniquify_algorithm", true}, {"best_effort_uniquify", false},
        {"uniquify_percent", true}, {"num_runs", true}, {"advance_direction", true}};
    std::map<char, std::string> shorts = {
        {'m', "market"}, {'d', "json_dir"}, {'f', "json_file"}, {'t', "tag"}, {'n', "num_runs"}};
    if (has_sources) {
      known["src"] = true;
      shorts['s'] = "src";
    }
    if (has_validate)
      known["validate"] = false;

    std::map<std::string, std::string> got;
    for (int i = 1; i < argc; ++i) {
      std::string a = argv[i], name, value;
      bool has_value = false;
      if (a.rfind("--", 0) == 0) {
        name = a.substr(2);
        std::size_t eq = name.find('=');
        if (eq != std::string::npos) {
          value = name.substr(eq + 1);
          name = name.substr(0, eq);
          has_value = true;
        }
      } else if (a.size() >= 2 && a[0] == '-' && shorts.count(a[1])) {
        name = shorts[a[1]];
        if (a.size() > 2) {
          value = a.substr(2);
          has_value = true;
        }
      } else {
        fail("unrecognised argument '" + a + "'");
      }
      if (!known.count(name))
        fail("unknown option '" + a + "'");
      if (known[name] && !has_value) {
        if (i + 1 >= argc)
          fail("option '" + a + "' needs a value");
        value = argv[++i];
      }
      got[name] = value;
    }

    if (got.count("help") || !got.count("market"))
      usage_and_exit(has_sources, has_validate);
    filename = got["market"];
    if (util::is_binary_csr(filename))
      binary = true;
    else if (!util::is_market(filename))
      usage_and_exit(has_sources, has_validate);

    validate = got.count("validate") > 0;
    export_metrics = got.count("export_metrics") > 0;
    if (got.count("num_runs"))
      num_runs = to_int(got["num_runs"], "num_runs");
    if (got.count("tag"))
      tag_string = got["tag"];
    if (got.count("src"))
      source_string = got["src"];
    if (got.count("json_dir"))
      json_dir = got["json_dir"];
    if (got.count("json_file"))
      json_file = got["json_file"];
    if (got.count("advance_load_balance"))
      advance_load_balance = parse_load_balance(got["advance_load_balance"]);
    if (got.count("filter_algorithm"))
      filter_algorithm = parse_filter_algorithm(got["filter_algorithm"]);
    enable_filter = got.count("enable_filter") > 0;
    enable_uniquify = got.count("enable_uniquify") > 0;
    if (got.count("uniquify_algorithm"))
      uniquify_algorithm = parse_uniquify_algorithm(got["uniquify_algorithm"]);
    if (got.count("best_effort_uniquify"))
      best_effort_uniquify = true;
    if (got.count("uniquify_percent")) {
      try {
        uniquify_percent = std::stof(got["uniquify_percent"]);
      } catch (...) {
        fail("uniquify_percent needs a number");
      }
    }
    if (got.count("advance_direction"))
      advance_direction = parse_advance_direction(got["advance_direction"]);
  }

  gunrock::options_t get_options() const {
    gunrock::options_t opts;
    opts.advance_load_balance = advance_load_balance;
    opts.filter_algorithm = filter_algorithm;
    opts.enable_filter = enable_filter;
    opts.enable_uniquify = enable_uniquify;
    opts.uniquify_algorithm = uniquify_algorithm;
    opts.best_effort_uniquify = best_effort_uniquify;
    opts.uniquify_percent = uniquify_percent;
    opts.advance_direction = advance_direction;
    return opts;
  }

 private:
  std::string program, description;

  static int to_int(const std::string& s, const char* what) {
    try {
      return std::stoi(s);
    } catch (...) {
      std::cerr << "Error: option " << what << " needs an integer" << std::endl;
      std::exit(1);
    }
  }
  void fail(const std::string& why) const {
    std::cerr << "Error parsing options: " << why << std::endl;
    std::exit(1);
  }
  void usage_and_exit(bool has_sources, bool has_validate) const {
    std::cout << description << "\nUsage:\n  " << program << " [OPTION...]\n\n"
              << "      --help                      Print help\n"
              << "      --export_metrics            export performance analysis metrics\n"
              << "  -m, --market arg                Matrix file (.mtx, or binary .csr)\n"
              << "  -d, --json_dir arg              JSON output directory\n"
              << "  -f, --json_file arg             JSON output file\n"
              << "  -t, --tag arg                   Tags for the JSON output; comma-separated\n"
              << "      --advance_load_balance arg  thread_mapped, block_mapped, merge_path, ...\n"
              << "      --advance_direction arg     forward, backward, optimized\n"
              << "      --filter_algorithm arg      remove, predicated, compact, bypass\n"
              << "      --enable_filter             Enable filter operator\n"
              << "      --enable_uniquify           Enable uniquify operator\n"
              << "      --uniquify_algorithm arg    unique, unique_copy\n"
              << "      --best_effort_uniquify      Best-effort uniquification (skip sorting)\n"
              << "      --uniquify_percent arg      Percentage of elements to uniquify (0-100)\n";
    if (has_sources)
      std::cout << "  -s, --src arg                   Source(s) (random if omitted); comma-separated\n"
                << "  -n, --num_runs arg              Number of runs (ignored if multiple sources)\n";
    else
      std::cout << "  -n, --num_runs arg              Number of runs\n";
    if (has_validate)
      std::cout << "      --validate                  CPU validation\n";
    std::cout << std::endl;
    std::exit(0);
  }
};

inline void parse_source_string(std::string source_str,
                                std::vector<int>* source_vect,
                                int n_vertices,
                                int n_runs) {
  if (source_str == "") {
    std::random_device seed;
    std::mt19937 engine(seed());
    std::uniform_int_distribution<int> dist(0, n_vertices - 1);
    for (int i = 0; i < n_runs; i++)
      source_vect->push_back(dist(engine));
    return;
  }
  std::stringstream ss(source_str);
  std::string token;
  while (std::getline(ss, token, ',')) {
    int v = -1;
    try {
      v = std::stoi(token);
    } catch (...) {
      v = -1;
    }
    if (v < 0 || v >= n_vertices) {
      std::cout << "Error: Invalid source\n";
      std::exit(1);
    }
    source_vect->push_back(v);
  }
  if (source_vect->size() == 1 && n_runs > 1)
    source_vect->insert(source_vect->end(), n_runs - 1, source_vect->at(0));
}

inline void parse_tag_string(std::string tag_str, std::vector<std::string>* tag_vect) {
  std::stringstream ss(tag_str);
  std::string tag;
  while (std::getline(ss, tag, ','))
    if (tag != "")
      tag_vect->push_back(tag);
}

}  // namespace cli
}  // namespace io
}  // namespace gunrock
