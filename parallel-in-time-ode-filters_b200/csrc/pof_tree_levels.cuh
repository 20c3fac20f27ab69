#pragma once

namespace pof {

// ------------------------------------------------------------------------------------------------------------
// Tree bookkeeping (host side): level l has sz[l] nodes, stored at node offset off[l]; level 0 = chunks.
// Node i of level l+1 has children 2i and 2i+1 of level l (the second may be missing).
// ------------------------------------------------------------------------------------------------------------
struct TreeLevels {
  static constexpr int MAXL = 48;
  int nlev;
  long sz[MAXL], off[MAXL], total;
  void build(long nchunks) {
    nlev = 0;
    total = 0;
    long s = nchunks;
    while (true) {
      sz[nlev] = s;
      off[nlev] = total;
      total += s;
      ++nlev;
      if (s <= 1) break;
      s = (s + 1) / 2;
    }
  }
};

}  // namespace pof
