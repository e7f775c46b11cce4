// TEST INFRASTRUCTURE ONLY (oracle/).
// The reference's src/matcher/array_allocator.h:170-172 declares its dynamic specialisation as
// `template <> template <typename Base> struct _Array2DAllocator<0,0,Base>`, which GCC >= 5
// rejects ("too many template-parameter-lists"). To compile the reference's gridmap.h /
// chargrid.cpp verbatim we pre-define that header's include guard and supply an equivalent
// dynamic rows x cols store here. Only the members gridmap.h uses are provided
// (gridmap.h:68,134,188,205: operator[], rows(), cols(), value semantics).
// Like the reference (array_allocator.h:243-245) every row is its own heap block, so
// cell(x,y) = rows[x][y] with y contiguous.
#ifndef CGM_ORACLE_REF_ARRAY_ALLOCATOR_FIX_H
#define CGM_ORACLE_REF_ARRAY_ALLOCATOR_FIX_H

#define _ARRAY_ALLOCATOR_HH_  // suppress the reference header body

#include <cassert>
#include <vector>

template <int Rows, int Cols, typename Base>
struct _Array2DAllocator;

template <typename Base>
struct _Array2DAllocator<0, 0, Base> {
  typedef Base BaseType;

  explicit _Array2DAllocator(int rows = 0, int cols = 0) { reshape(rows, cols); }

  Base* operator[](int r) { return store_[r].data(); }
  const Base* operator[](int r) const { return store_[r].data(); }

  int rows() const { return rows_; }
  int cols() const { return cols_; }

  void resize(int rows, int cols) {
    if (rows == rows_ && cols == cols_) return;
    reshape(rows, cols);
  }

 private:
  void reshape(int rows, int cols) {
    if (rows <= 0 || cols <= 0) {  // reference: a zero extent yields an empty 0x0 store
      rows_ = cols_ = 0;
      store_.clear();
      return;
    }
    rows_ = rows;
    cols_ = cols;
    store_.assign(rows, std::vector<Base>(cols));
  }

  std::vector<std::vector<Base> > store_;
  int rows_ = 0;
  int cols_ = 0;
};

#endif
