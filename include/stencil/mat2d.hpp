#pragma once
// Mat2D<T>: small dense row-major matrix (rank-to-rank byte matrices, bandwidth / distance matrices
// of the placement QAP).  shape().x = columns, shape().y = rows; at(i, j) = row i, column j.

#include <algorithm>
#include <cassert>
#include <cstdint>
#include <cstring>
#include <initializer_list>
#include <iostream>
#include <limits>
#include <vector>

struct Shape {
  uint64_t x;
  uint64_t y;
  Shape(uint64_t x_, uint64_t y_) : x(x_), y(y_) {}
  uint64_t flatten() const noexcept { return x * y; }
  bool operator==(const Shape &o) const noexcept { return x == o.x && y == o.y; }
  bool operator!=(const Shape &o) const noexcept { return !(*this == o); }
};

template <typename T> class Mat2D {
public:
  // a mutable view of one row
  class Row {
    friend class Mat2D;
    T *p_;
    int64_t n_;
    Row(T *p, int64_t n) : p_(p), n_(n) {}

  public:
    T &operator[](int64_t i) noexcept { return p_[i]; }
    const T &operator[](int64_t i) const noexcept { return p_[i]; }
    T *begin() const { return p_; }
    T *end() const { return p_ + n_; }
    Row &operator=(const Row &o) {
      assert(o.n_ == n_);
      std::copy(o.p_, o.p_ + n_, p_);
      return *this;
    }
    Row &operator=(const std::vector<T> &o) {
      assert(int64_t(o.size()) == n_);
      std::copy(o.begin(), o.end(), p_);
      return *this;
    }
  };
  class ConstRow {
    friend class Mat2D;
    const T *p_;
    int64_t n_;
    ConstRow(const T *p, int64_t n) : p_(p), n_(n) {}

  public:
    const T &operator[](int64_t i) const noexcept { return p_[i]; }
    const T *begin() const { return p_; }
    const T *end() const { return p_ + n_; }
  };

  std::vector<T> data_;
  Shape shape_;

  Mat2D() : shape_(0, 0) {}
  Mat2D(int64_t x, int64_t y) : data_(size_t(x * y)), shape_(x, y) {}
  Mat2D(int64_t x, int64_t y, const T &v) : data_(size_t(x * y), v), shape_(x, y) {}
  Mat2D(Shape s) : Mat2D(int64_t(s.x), int64_t(s.y)) {}
  Mat2D(Shape s, const T &v) : Mat2D(int64_t(s.x), int64_t(s.y), v) {}
  Mat2D(const std::initializer_list<std::initializer_list<T>> &rows) : shape_(0, 0) {
    for (const auto &r : rows) push_back(std::vector<T>(r));
  }
  Mat2D(const Mat2D &) = default;
  Mat2D(Mat2D &&) = default;
  Mat2D &operator=(const Mat2D &) = default;
  Mat2D &operator=(Mat2D &&) = default;

  T &at(int64_t i, int64_t j) noexcept {
    assert(i < int64_t(shape_.y) && j < int64_t(shape_.x));
    return data_[size_t(i) * shape_.x + size_t(j)];
  }
  const T &at(int64_t i, int64_t j) const noexcept {
    assert(i < int64_t(shape_.y) && j < int64_t(shape_.x));
    return data_[size_t(i) * shape_.x + size_t(j)];
  }
  Row operator[](int64_t i) noexcept { return Row(data_.data() + size_t(i) * shape_.x, int64_t(shape_.x)); }
  ConstRow operator[](int64_t i) const noexcept { return ConstRow(data_.data() + size_t(i) * shape_.x, int64_t(shape_.x)); }

  T *data() noexcept { return data_.data(); }
  const T *data() const noexcept { return data_.data(); }

  // grow or shrink to x columns, y rows keeping the top-left block
  void resize(int64_t x, int64_t y) {
    Mat2D next(x, y);
    const int64_t rows = std::min<int64_t>(y, int64_t(shape_.y)), cols = std::min<int64_t>(x, int64_t(shape_.x));
    for (int64_t i = 0; i < rows; ++i)
      for (int64_t j = 0; j < cols; ++j) next.at(i, j) = at(i, j);
    data_.swap(next.data_);
    shape_ = next.shape_;
  }

  void push_back(const std::vector<T> &row) {
    assert(shape_.y == 0 || row.size() == shape_.x);
    resize(int64_t(row.size()), int64_t(shape_.y) + 1);
    (*this)[int64_t(shape_.y) - 1] = row;
  }

  const Shape &shape() const noexcept { return shape_; }
  uint64_t size() const noexcept { return shape_.flatten(); }

  bool operator==(const Mat2D &o) const noexcept { return shape_ == o.shape_ && data_ == o.data_; }

  template <typename S> Mat2D &operator/=(const S &s) {
    for (auto &e : data_) e /= s;
    return *this;
  }
};

// element-wise 1/x with 1/0 = +inf (bandwidth matrix -> distance matrix)
inline Mat2D<double> make_reciprocal(const Mat2D<double> &m) {
  Mat2D<double> out(m.shape());
  for (uint64_t k = 0; k < m.size(); ++k) {
    const double e = m.data()[k];
    out.data()[k] = (0 == e) ? std::numeric_limits<double>::infinity() : 1.0 / e;
  }
  return out;
}
