// tinymatrix.hpp -- fixed-size row-major matrix with a name (interface of reference
// include/tinymatrix.hpp:41-307: mat(i,j), mat[i][j], numRows/numCols/size, getPtr, ==, !=,
// getSubMatrix, print, tostr, public `name`).  At the C ABI a TM2x2 is flattened to double[8].
#ifndef IQS_TINYMATRIX_HPP
#define IQS_TINYMATRIX_HPP
#include <cassert>
#include <cstdio>
#include <initializer_list>
#include <iostream>
#include <string>

namespace iqs {

template <class ValueType, unsigned M, unsigned N = M, unsigned align = alignof(ValueType)>
class TinyMatrix {
 public:
  using value_type = ValueType;
  using pointer = ValueType *;
  using const_pointer = ValueType const *;
  using reference = ValueType &;
  using size_type = unsigned;
  using RowType = ValueType[N];

  TinyMatrix() { static_assert(N * M != 0, "A zero-dimensional matrix is not allowed."); }

  template <class U>
  TinyMatrix(U init[M][N]) {
    for (size_type i = 0; i < M; ++i)
      for (size_type j = 0; j < N; ++j) data_[i][j] = init[i][j];
  }
  template <class U>
  TinyMatrix(std::initializer_list<std::initializer_list<U>> const &init) {
    size_type i = 0;
    for (auto const &row : init) {
      size_type j = 0;
      for (auto const &e : row) data_[i][j++] = e;
      ++i;
    }
  }
  template <class U, unsigned alignrhs>
  TinyMatrix(TinyMatrix<U, M, N, alignrhs> const &rhs) {
    for (size_type i = 0; i < M; ++i)
      for (size_type j = 0; j < N; ++j) data_[i][j] = rhs(i, j);
  }
  TinyMatrix(TinyMatrix const &) = default;
  TinyMatrix &operator=(TinyMatrix const &) = default;
  template <class U, unsigned alignrhs>
  TinyMatrix &operator=(TinyMatrix<U, M, N, alignrhs> const &rhs) {
    for (size_type i = 0; i < M; ++i)
      for (size_type j = 0; j < N; ++j) data_[i][j] = rhs(i, j);
    return *this;
  }
  template <class U>
  TinyMatrix &operator=(U const (&rhs)[M][N]) {
    for (size_type i = 0; i < M; ++i)
      for (size_type j = 0; j < N; ++j) data_[i][j] = rhs[i][j];
    return *this;
  }

  constexpr size_type numRows() const { return M; }
  constexpr size_type numCols() const { return N; }
  constexpr size_type size() const { return N * M; }

  value_type operator()(size_type i, size_type j) const {
    assert(i < M && "Row index out of range");
    assert(j < N && "Column index out of range");
    return data_[i][j];
  }
  reference operator()(size_type i, size_type j) {
    assert(i < M && "Row index out of range");
    assert(j < N && "Column index out of range");
    return data_[i][j];
  }

  template <class U, unsigned alignrhs>
  bool operator==(TinyMatrix<U, M, N, alignrhs> const &rhs) const {
    for (size_type i = 0; i < M; ++i)
      for (size_type j = 0; j < N; ++j)
        if (data_[i][j] != rhs(i, j)) return false;
    return true;
  }
  template <class U, unsigned alignrhs>
  bool operator!=(TinyMatrix<U, M, N, alignrhs> const &rhs) const { return !(*this == rhs); }
  template <class U>
  bool operator==(U const (&rhs)[M][N]) {
    for (size_type i = 0; i < M; ++i)
      for (size_type j = 0; j < N; ++j)
        if (data_[i][j] != rhs[i][j]) return false;
    return true;
  }
  template <class U>
  bool operator!=(U const (&rhs)[M][N]) { return !(*this == rhs); }

  const_pointer getPtr() const { return &data_[0][0]; }
  RowType &operator[](unsigned i) { assert(i < M); return data_[i]; }
  RowType const &operator[](unsigned i) const { assert(i < M); return data_[i]; }

  template <unsigned MSub, unsigned NSub = MSub>
  TinyMatrix<ValueType, MSub, NSub, align> getSubMatrix(unsigned i_start = 0, unsigned j_start = 0,
                                                        unsigned i_stride = 1, unsigned j_stride = 1) const {
    assert(i_stride > 0 && j_stride > 0);
    assert((MSub - 1) * i_stride + i_start < M && (NSub - 1) * j_stride + j_start < N);
    TinyMatrix<ValueType, MSub, NSub, align> sub;
    for (unsigned a = 0; a < MSub; ++a)
      for (unsigned b = 0; b < NSub; ++b) sub(a, b) = (*this)(i_start + a * i_stride, j_start + b * j_stride);
    return sub;
  }

  void print(std::string label) {
    printf("name: %s\n", label.c_str());
    for (size_type i = 0; i < M; ++i) {
      for (size_type j = 0; j < N; ++j) std::cout << data_[i][j].real() << " + i*" << data_[i][j].imag() << " ";
      printf("\n");
    }
  }
  std::string tostr() const {
    std::string s = "{";
    char buf[128];
    for (size_type i = 0; i < M; ++i) {
      s += "{";
      for (size_type j = 0; j < N; ++j) {
        snprintf(buf, sizeof(buf), "%.3lf+%.3lf ", (double)data_[i][j].real(), (double)data_[i][j].imag());
        s += buf;
      }
      s += "}";
    }
    return s + "{";
  }

  std::string name;

 protected:
  alignas(align == 0 ? 8 : align) ValueType data_[M][N];
};

}  // namespace iqs
#endif
