#pragma once
// DirectionMap<T>: one T per direction vector in {-1,0,1}^3, stored [z+1][y+1][x+1].

#include <array>
#include <cassert>

template <typename T> class DirectionMap {
  std::array<T, 27> v_{};

  static constexpr int slot(int x, int y, int z) { return z * 9 + y * 3 + x; }

public:
  typedef int index_type;

  DirectionMap() = default;
  explicit DirectionMap(const T &fill) { v_.fill(fill); }

  bool operator==(const DirectionMap &o) const noexcept { return v_ == o.v_; }

  // index by 0..2 per axis
  T &at(index_type x, index_type y, index_type z) noexcept {
    assert(x >= 0 && x <= 2 && y >= 0 && y <= 2 && z >= 0 && z <= 2);
    return v_[slot(x, y, z)];
  }
  const T &at(index_type x, index_type y, index_type z) const noexcept {
    assert(x >= 0 && x <= 2 && y >= 0 && y <= 2 && z >= 0 && z <= 2);
    return v_[slot(x, y, z)];
  }
  // index by direction component -1..1 per axis
  T &at_dir(index_type x, index_type y, index_type z) noexcept { return at(x + 1, y + 1, z + 1); }
  const T &at_dir(index_type x, index_type y, index_type z) const noexcept { return at(x + 1, y + 1, z + 1); }

  const T *data() const noexcept { return v_.data(); }
};
