#pragma once

namespace gpu_topo {

// relative bandwidth between two CUDA devices of this node (bigger = faster; only ratios matter)
double bandwidth(int src, int dst);

// try to give `src` peer access to `dst` (memoised)
void enable_peer(const int src, const int dst);

// can `src` address `dst`'s memory?
bool peer(const int src, const int dst);

} // namespace gpu_topo
