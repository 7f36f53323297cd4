// matrix_alloc_jpd.hpp -- drop-in for the reference's 2-D array helper
// (/root/reference/include/matrix_alloc_jpd.hpp:19-25: Alloc2D / Dealloc2D), needed because
// callers such as the reference's tests/test_jacobi.cpp (:20, :270-271, :479-481) include it
// next to jacobi_pd.hpp.  Same contract -- a row-pointer table over ONE contiguous row-major
// block, `(*p)[i][j]` indexing, Dealloc2D frees it and nulls the pointer -- written
// independently: here the table and the payload share a single heap block.
#ifndef JPD_B200_MATRIX_ALLOC_HPP_
#define JPD_B200_MATRIX_ALLOC_HPP_

#include <cstddef>
#include <cstdint>
#include <new>

namespace matrix_alloc_jpd {

namespace detail {
// bytes reserved for the row table, rounded so that the payload keeps max alignment
inline std::size_t table_bytes(std::size_t nrows, std::size_t ptr_size) {
  const std::size_t a = alignof(std::max_align_t);
  return ((nrows ? nrows : 1) * ptr_size + a - 1) / a * a;
}
}  // namespace detail

/// Allocate an nrows x ncols array addressable as (*paaX)[i][j] (row-major, contiguous).
template <typename Entry>
void Alloc2D(std::size_t nrows, std::size_t ncols, Entry ***paaX) {
  if (!paaX) return;
  const std::size_t head = detail::table_bytes(nrows, sizeof(Entry *));
  const std::size_t count = nrows * ncols;
  unsigned char *block = static_cast<unsigned char *>(::operator new(head + (count ? count : 1) * sizeof(Entry)));
  Entry **rows = reinterpret_cast<Entry **>(block);
  Entry *cells = reinterpret_cast<Entry *>(block + head);
  for (std::size_t k = 0; k < count; ++k) new (cells + k) Entry();
  rows[0] = cells;  // also valid for nrows == 0: Dealloc2D only needs the block start
  for (std::size_t r = 1; r < nrows; ++r) rows[r] = cells + r * ncols;
  *paaX = rows;
}

/// Release an array made by Alloc2D and set the caller's pointer to nullptr.
template <typename Entry>
void Dealloc2D(Entry ***paaX) {
  if (!paaX || !*paaX) return;
  ::operator delete(static_cast<void *>(*paaX));
  *paaX = nullptr;
}

}  // namespace matrix_alloc_jpd

#endif  // JPD_B200_MATRIX_ALLOC_HPP_
