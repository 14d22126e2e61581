/**
 * @file array.hxx
 * @brief `gunrock::array<T, N>`: a fixed-size aggregate usable from host and device code, with the
 * surface of the reference's std::array clone (include/gunrock/container/array.hxx:92-169: `data`, `size`,
 * `max_size`, `empty`, `operator[]`, the iterator typedefs) plus what std::array adds on top
 * (`begin`/`end`, `front`/`back`, `fill`, comparisons).  Off the traversal path: launch dimensions and small
 * per-thread tuples.  Zero-length arrays are allowed and hold no storage that is ever dereferenced.
 */
#pragma once

#include <cstddef>
#include <iterator>

#ifndef __host__
#define GUNROCK_B200_ARRAY_UNDEF_HD 1
#define __host__
#define __device__
#endif

namespace gunrock {

template <typename T, std::size_t NumElements>
struct array {
  using value_type = T;
  using pointer_t = value_type*;
  using const_pointer_t = const value_type*;
  using reference_t = value_type&;
  using const_reference_t = const value_type&;
  using iterator = value_type*;
  using const_iterator = const value_type*;
  using size_type = std::size_t;
  using difference_type = std::ptrdiff_t;
  using reverse_iterator = std::reverse_iterator<iterator>;
  using const_reverse_iterator = std::reverse_iterator<const_iterator>;

  /// Public so that `array<int, 3> a = {1, 2, 3};` stays aggregate initialisation.
  T _elements[NumElements ? NumElements : 1];

  __host__ __device__ constexpr pointer_t data() noexcept { return _elements; }
  __host__ __device__ constexpr const_pointer_t data() const noexcept { return _elements; }
  __host__ __device__ constexpr size_type size() const noexcept { return NumElements; }
  __host__ __device__ constexpr size_type max_size() const noexcept { return NumElements; }
  __host__ __device__ constexpr bool empty() const noexcept { return NumElements == 0; }

  __host__ __device__ constexpr reference_t operator[](size_type n) noexcept { return _elements[n]; }
  __host__ __device__ constexpr const_reference_t operator[](size_type n) const noexcept { return _elements[n]; }
  __host__ __device__ constexpr reference_t front() noexcept { return _elements[0]; }
  __host__ __device__ constexpr const_reference_t front() const noexcept { return _elements[0]; }
  __host__ __device__ constexpr reference_t back() noexcept { return _elements[NumElements ? NumElements - 1 : 0]; }
  __host__ __device__ constexpr const_reference_t back() const noexcept {
    return _elements[NumElements ? NumElements - 1 : 0];
  }

  __host__ __device__ constexpr iterator begin() noexcept { return _elements; }
  __host__ __device__ constexpr const_iterator begin() const noexcept { return _elements; }
  __host__ __device__ constexpr const_iterator cbegin() const noexcept { return _elements; }
  __host__ __device__ constexpr iterator end() noexcept { return _elements + NumElements; }
  __host__ __device__ constexpr const_iterator end() const noexcept { return _elements + NumElements; }
  __host__ __device__ constexpr const_iterator cend() const noexcept { return _elements + NumElements; }

  __host__ __device__ constexpr void fill(const T& value) {
    for (size_type i = 0; i < NumElements; ++i)
      _elements[i] = value;
  }
};

template <typename T, std::size_t N>
__host__ __device__ constexpr bool operator==(const array<T, N>& a, const array<T, N>& b) {
  for (std::size_t i = 0; i < N; ++i)
    if (!(a[i] == b[i]))
      return false;
  return true;
}
template <typename T, std::size_t N>
__host__ __device__ constexpr bool operator!=(const array<T, N>& a, const array<T, N>& b) {
  return !(a == b);
}

}  // namespace gunrock

#ifdef GUNROCK_B200_ARRAY_UNDEF_HD
#undef __host__
#undef __device__
#undef GUNROCK_B200_ARRAY_UNDEF_HD
#endif
