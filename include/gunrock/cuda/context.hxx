/**
 * @file context.hxx
 * @brief `gcuda::standard_context_t` / `gcuda::multi_context_t`
 * (include/gunrock/cuda/context.hxx:54-216): one non-blocking stream + event + timer + device
 * properties per device, `execution_policy()` for Thrust boundary calls, peer access.
 * B200 addition: every context owns the operator workspace (control blocks, hub list, scan
 * scratch) so no operator launch ever calls cudaMalloc (include/gunrock/b200/runtime.cuh).
 */
#pragma once

#include <memory>
#include <typeindex>
#include <typeinfo>
#include <vector>

#include <cuda_runtime.h>
#include <thrust/execution_policy.h>
#include <thrust/host_vector.h>
#include <thrust/system/cuda/execution_policy.h>

#include <gunrock/b200/runtime.cuh>
#include <gunrock/error.hxx>
#include <gunrock/util/timer.hxx>

namespace gunrock {
namespace gcuda {

typedef int device_id_t;
typedef cudaStream_t stream_t;
typedef cudaEvent_t event_t;
typedef cudaDeviceProp device_properties_t;

struct context_t {
  context_t() = default;
  context_t(const context_t&) = delete;
  context_t& operator=(const context_t&) = delete;
  virtual ~context_t() = default;
  virtual const device_properties_t& props() const = 0;
  virtual void print_properties() = 0;
  virtual stream_t stream() = 0;
  virtual void synchronize() = 0;
  virtual event_t event() = 0;
  virtual util::timer_t& timer() = 0;
};

class standard_context_t : public context_t {
 protected:
  device_properties_t _props;
  device_id_t _ordinal;
  stream_t _stream = nullptr;
  event_t _event = nullptr;
  util::timer_t _timer;
  b200::workspace_t _workspace;
  bool _owns_stream = false;
#ifdef GUNROCK_B200_REFERENCE_ADVANCE_OUTPUT
  bool _reference_advance_output = true;
#else
  bool _reference_advance_output = false;
#endif
  /// Scratch of the fused enactors (bfs / sssp / pr ...), one object per type, created on first use and
  /// destroyed WITH the context (device buffers, pinned blocks, events).  Nothing outside the context
  /// keeps a pointer-keyed table of contexts, so a freed-and-reallocated context can never alias
  /// another one's scratch.
  struct slot_t {
    std::type_index type;
    std::shared_ptr<void> object;
  };
  std::vector<slot_t> _scratch;

 public:
  standard_context_t(device_id_t device = 0) : context_t(), _ordinal(device) {
    error::throw_if_exception(cudaSetDevice(_ordinal), "cudaSetDevice");
    error::throw_if_exception(cudaGetDeviceProperties(&_props, _ordinal), "cudaGetDeviceProperties");
    error::throw_if_exception(cudaStreamCreateWithFlags(&_stream, cudaStreamNonBlocking),
                              "cudaStreamCreate");
    _owns_stream = true;
    error::throw_if_exception(cudaEventCreateWithFlags(&_event, cudaEventDisableTiming),
                              "cudaEventCreate");
    _workspace.init(_stream);
  }
  standard_context_t(cudaStream_t stream, device_id_t device = 0)
      : context_t(), _ordinal(device), _stream(stream) {
    error::throw_if_exception(cudaSetDevice(_ordinal), "cudaSetDevice");
    error::throw_if_exception(cudaGetDeviceProperties(&_props, _ordinal), "cudaGetDeviceProperties");
    error::throw_if_exception(cudaEventCreateWithFlags(&_event, cudaEventDisableTiming),
                              "cudaEventCreate");
    _workspace.init(_stream);
  }
  ~standard_context_t() {
    int before = 0;
    cudaGetDevice(&before);
    cudaSetDevice(_ordinal);
    if (_stream)
      cudaStreamSynchronize(_stream);  // nothing of ours is still in flight when the scratch goes
    _scratch.clear();
    cudaEventDestroy(_event);
    if (_owns_stream && _stream)
      cudaStreamDestroy(_stream);      // only a stream this context created itself
    cudaSetDevice(before);
  }

  const device_properties_t& props() const override { return _props; }
  void print_properties() override {
    std::cout << _props.name << " : sm_" << _props.major << _props.minor << ", "
              << _props.multiProcessorCount << " SMs, "
              << (_props.totalGlobalMem >> 20) << " MiB" << std::endl;
  }
  stream_t stream() override { return _stream; }
  void synchronize() override {
    error::throw_if_exception(
        _stream ? cudaStreamSynchronize(_stream) : cudaDeviceSynchronize(), "synchronize");
  }
  event_t event() override { return _event; }
  util::timer_t& timer() override { return _timer; }
  device_id_t ordinal() { return _ordinal; }
  auto execution_policy() { return thrust::cuda::par_nosync.on(_stream); }
  /// B200 operator scratch bound to this context's stream.
  b200::workspace_t& workspace() { return _workspace; }
  /// Opt-in compatibility switch for `operators::advance::execute` (default off; on by default when the TU is
  /// compiled with -DGUNROCK_B200_REFERENCE_ADVANCE_OUTPUT).  Off: the output frontier is compact.  On: it has
  /// the reference's shape -- one slot per (input entry, out-edge) in edge-rank order, -1 where the functor
  /// returned false, size = the input frontier's out-degree sum (merge_path.hxx:218-279) -- for code that
  /// indexes an advance's output by edge rank or counts its invalid slots.
  bool reference_advance_output() const { return _reference_advance_output; }
  void reference_advance_output(bool on) { _reference_advance_output = on; }
  /// The context-owned scratch object of type T (default-constructed on first use).
  template <typename T>
  T& scratch() {
    const std::type_index key(typeid(T));
    for (auto& s : _scratch)
      if (s.type == key)
        return *static_cast<T*>(s.object.get());
    _scratch.push_back({key, std::shared_ptr<void>(new T(), [](void* p) { delete static_cast<T*>(p); })});
    return *static_cast<T*>(_scratch.back().object.get());
  }
};

class multi_context_t {
 public:
  thrust::host_vector<standard_context_t*> contexts;
  thrust::host_vector<device_id_t> devices;
  static constexpr std::size_t MAX_NUMBER_OF_GPUS = 1024;

  multi_context_t(thrust::host_vector<device_id_t> _devices) : devices(_devices) {
    for (auto& d : devices)
      contexts.push_back(new standard_context_t(d));
  }
  multi_context_t(thrust::host_vector<device_id_t> _devices, cudaStream_t _stream)
      : devices(_devices) {
    for (auto& d : devices)
      contexts.push_back(new standard_context_t(_stream, d));
  }
  multi_context_t(device_id_t _device) : devices(1, _device) {
    contexts.push_back(new standard_context_t(_device));
  }
  multi_context_t(device_id_t _device, cudaStream_t _stream) : devices(1, _device) {
    contexts.push_back(new standard_context_t(_stream, _device));
  }
  multi_context_t(const multi_context_t&) = delete;
  multi_context_t& operator=(const multi_context_t&) = delete;
  ~multi_context_t() {
    for (auto& c : contexts)
      delete c;
  }

  auto get_context(device_id_t device) {
    auto* c = contexts[device];
    cudaSetDevice(c->ordinal());
    return c;
  }
  auto size() { return contexts.size(); }

  void enable_peer_access() {
    int n = static_cast<int>(size());
    for (int i = 0; i < n; ++i) {
      cudaSetDevice(devices[i]);
      for (int j = 0; j < n; ++j) {
        if (i == j)
          continue;
        int can = 0;
        cudaDeviceCanAccessPeer(&can, devices[i], devices[j]);
        if (can) {
          cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
          if (e == cudaErrorPeerAccessAlreadyEnabled)
            cudaGetLastError();
          else
            error::throw_if_exception(e, "cudaDeviceEnablePeerAccess");
        }
      }
    }
    cudaSetDevice(devices[0]);
  }
};

}  // namespace gcuda
}  // namespace gunrock
