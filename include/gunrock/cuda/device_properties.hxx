/**
 * @file device_properties.hxx
 * @brief `gcuda::compute_capability_t` and the `gcuda::properties::` tables keyed on it
 * (include/gunrock/cuda/device_properties.hxx:21-300): compile-time limits of an SM by compute capability, and
 * run-time queries on a `device_properties_t`.  Host-only constexpr code; the B200 kernels size themselves from
 * `b200::device_info_t` at run time and do not use it.
 *
 * Beyond the reference, whose tables stop at sm_90: 10.0 / 10.3 (B200 / B300: 2048 threads, 32 CTAs, 64 K
 * registers, 228 KiB shared memory per SM) and 12.x (1536 threads, 24 CTAs, 100 KiB).
 */
#pragma once

#include <cstddef>
#include <string>

#include <cuda_runtime.h>

namespace gunrock {
namespace gcuda {

#ifndef GUNROCK_B200_DEVICE_TYPEDEFS
#define GUNROCK_B200_DEVICE_TYPEDEFS
typedef int device_id_t;
typedef cudaDeviceProp device_properties_t;
#endif

struct compute_capability_t {
  unsigned major;
  unsigned minor;
  constexpr unsigned as_combined_number() const { return major * 10 + minor; }
  constexpr bool operator==(int i) const { return static_cast<int>(as_combined_number()) == i; }
  constexpr bool operator!=(int i) const { return static_cast<int>(as_combined_number()) != i; }
  constexpr bool operator>(int i) const { return static_cast<int>(as_combined_number()) > i; }
  constexpr bool operator<(int i) const { return static_cast<int>(as_combined_number()) < i; }
  constexpr bool operator>=(int i) const { return static_cast<int>(as_combined_number()) >= i; }
  constexpr bool operator<=(int i) const { return static_cast<int>(as_combined_number()) <= i; }
};

constexpr compute_capability_t make_compute_capability(unsigned major, unsigned minor) {
  return compute_capability_t{major, minor};
}
/// 86 -> 8.6, 100 -> 10.0, 103 -> 10.3, 120 -> 12.0
constexpr compute_capability_t make_compute_capability(unsigned combined) {
  return compute_capability_t{combined / 10, combined % 10};
}

namespace properties {

enum : std::size_t { KiB = 1024, K = 1024 };

inline constexpr const char* arch_name(compute_capability_t cc) {
  return cc.major == 12 || cc.major == 10 ? "Blackwell"
         : cc == 90                       ? "Hopper"
         : cc == 89                       ? "Ada"
         : cc.major == 8                  ? "Ampere"
         : cc == 75                       ? "Turing"
         : cc.major == 7                  ? "Volta"
         : cc.major == 6                  ? "Pascal"
         : cc.major == 5                  ? "Maxwell"
         : cc.major == 3                  ? "Kepler"
                                          : nullptr;
}

inline constexpr unsigned cta_max_threads() { return 1024; }
inline constexpr unsigned warp_max_threads() { return 32; }

/// Resident thread blocks per SM.
inline constexpr unsigned sm_max_ctas(compute_capability_t cc) {
  return cc.major == 12                       ? 24
         : cc.major >= 9                      ? 32
         : cc == 89                           ? 24
         : (cc == 86 || cc == 87 || cc == 75) ? 16
         : cc.major >= 5                      ? 32
                                              : 16;
}
/// Resident threads per SM.
inline constexpr unsigned sm_max_threads(compute_capability_t cc) {
  return cc.major == 12                         ? 1536
         : cc.major >= 9                        ? 2048
         : (cc == 86 || cc == 87 || cc == 89)   ? 1536
         : cc == 75                             ? 1024
                                                : 2048;
}
/// 32-bit registers per SM.
inline constexpr unsigned sm_registers(compute_capability_t cc) {
  return cc == 37 ? 128 * K : 64 * K;
}
/// Shared memory an SM can be configured with, bytes.
inline constexpr unsigned sm_max_shared_memory_bytes(compute_capability_t cc) {
  return cc.major == 12                 ? 100 * KiB
         : cc.major >= 9                ? 228 * KiB
         : cc == 80 || cc == 87         ? 164 * KiB
         : cc == 86 || cc == 89         ? 100 * KiB
         : cc == 75                     ? 64 * KiB
         : cc.major == 7                ? 96 * KiB
         : cc == 61 || cc == 62         ? (cc == 61 ? 96 * KiB : 64 * KiB)
         : cc == 60                     ? 64 * KiB
         : cc == 52                     ? 96 * KiB
         : cc.major == 5                ? 64 * KiB
         : cc == 37                     ? 112 * KiB
                                        : 48 * KiB;
}
inline constexpr unsigned shared_memory_banks() { return 32; }
inline constexpr unsigned shared_memory_bank_stride() { return 4; }

// ---- run-time queries on a device_properties_t -------------------------------------------------------------
inline unsigned clock_rate(device_properties_t&, device_id_t device = 0) {
  int khz = 0;  // cudaDeviceProp::clockRate is gone from recent toolkits; the attribute stays
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
  return static_cast<unsigned>(khz);
}
inline unsigned compute_version(device_properties_t& prop) {
  return static_cast<unsigned>(prop.major * 10 + prop.minor);
}
inline std::string gpu_name(device_properties_t& prop) { return prop.name; }
inline unsigned sm_major(device_properties_t& prop) { return static_cast<unsigned>(prop.major); }
inline unsigned sm_minor(device_properties_t& prop) { return static_cast<unsigned>(prop.minor); }
inline unsigned multi_processor_count(device_properties_t& prop) {
  return static_cast<unsigned>(prop.multiProcessorCount);
}
inline std::size_t total_global_memory(device_properties_t& prop) { return prop.totalGlobalMem; }
inline int get_max_grid_dimension_x(device_id_t device) {
  int v = 0;
  cudaDeviceGetAttribute(&v, cudaDevAttrMaxGridDimX, device);
  return v;
}

}  // namespace properties
}  // namespace gcuda
}  // namespace gunrock
