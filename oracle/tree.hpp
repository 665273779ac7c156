// TEST INFRASTRUCTURE ONLY -- CPU oracle ("parity unpinned", see basis.hpp header).
//
// TreeMesh (hypercube, 2^d-tree, depth-first cell order) and the DG containers the reference builds by
// calling Trixi's CPU `init_elements / init_interfaces / init_boundaries / init_mortars`:
//   /root/reference/src/solvers/cache.jl:130-158   (3D; :60-88 2D; :15-38 1D)
//   /root/reference/src/solvers/containers_3d.jl:6-28,57-75,99-128,160-191 (field layouts)
// Conventions restated from Trixi.jl. Index arrays are 1-based like Julia.
#pragma once
#include <cstdint>
#include <vector>
#include <array>
#include <cmath>
#include <stdexcept>
#include "basis.hpp"

namespace orc {

struct RefinementBox { double lo[3], hi[3]; };

struct Cell {
  int level;
  int parent;               // -1 for root
  int child[8];             // -1 if none
  long ic[3];               // integer coordinates at `level`
  double x[3];              // centre
};

struct Tree {
  int ndim = 3, nchild = 8;
  double center0[3], length0 = 0;
  bool periodic[3] = {true, true, true};
  std::vector<Cell> cells;  // construction order (NOT depth-first); see dfs_order()

  bool has_children(int c) const { return cells[c].child[0] >= 0; }

  void init(int ndim_, const double* cmin, const double* cmax, const bool* per) {
    ndim = ndim_; nchild = 1 << ndim;
    length0 = 0;
    for (int d = 0; d < 3; ++d) { center0[d] = 0; periodic[d] = true; }
    for (int d = 0; d < ndim; ++d) {
      center0[d] = (cmin[d] + cmax[d]) / 2;
      length0 = std::max(length0, cmax[d] - cmin[d]);
      periodic[d] = per[d];
    }
    Cell root{};
    root.level = 0; root.parent = -1;
    for (int k = 0; k < 8; ++k) root.child[k] = -1;
    for (int d = 0; d < 3; ++d) { root.ic[d] = 0; root.x[d] = center0[d]; }
    cells.clear();
    cells.push_back(root);
  }

  double length_at_level(int level) const { return length0 / double(1L << level); }

  void refine_cell(int c) {
    if (has_children(c)) return;
    int lvl = cells[c].level + 1;
    double dx = length_at_level(lvl);
    for (int k = 0; k < nchild; ++k) {
      Cell ch{};
      ch.level = lvl; ch.parent = c;
      for (int q = 0; q < 8; ++q) ch.child[q] = -1;
      for (int d = 0; d < 3; ++d) { ch.ic[d] = 0; ch.x[d] = cells[c].x[d]; }
      for (int d = 0; d < ndim; ++d) {
        int bit = (k >> d) & 1;  // child k (0-based): + side in dim d iff bit d set
        ch.ic[d] = 2 * cells[c].ic[d] + bit;
        ch.x[d] = cells[c].x[d] + (bit ? 1.0 : -1.0) * dx / 2;
      }
      cells[c].child[k] = (int)cells.size();
      cells.push_back(ch);
    }
  }

  // Cell at (level, ic) or the coarser leaf covering it. Returns id and sets `exact` iff level matches.
  int find(int level, const long* ic, bool& exact) const {
    int c = 0;
    for (int l = 1; l <= level; ++l) {
      if (!has_children(c)) { exact = false; return c; }
      int k = 0;
      for (int d = 0; d < ndim; ++d) k |= (int)((ic[d] >> (level - l)) & 1) << d;
      c = cells[c].child[k];
    }
    exact = true;
    return c;
  }

  // Same-level neighbour in `direction` (1-based: 1 -x, 2 +x, 3 -y, ...). -1 if none at this level.
  // `coarse` is set when a coarser leaf covers the neighbour position (Trixi `has_coarse_neighbor`).
  int neighbor(int c, int direction, bool& coarse) const {
    coarse = false;
    int d = (direction - 1) / 2;
    long n = 1L << cells[c].level;
    long ic[3] = {cells[c].ic[0], cells[c].ic[1], cells[c].ic[2]};
    ic[d] += (direction % 2 == 0) ? 1 : -1;
    if (ic[d] < 0 || ic[d] >= n) {
      if (!periodic[d]) return -1;
      ic[d] = (ic[d] + n) % n;
    }
    bool exact;
    int nb = find(cells[c].level, ic, exact);
    if (!exact) { coarse = true; return -1; }
    return nb;
  }

  std::vector<int> leaves_dfs() const {
    std::vector<int> out, stack{0};
    while (!stack.empty()) {
      int c = stack.back(); stack.pop_back();
      if (!has_children(c)) { out.push_back(c); continue; }
      for (int k = nchild - 1; k >= 0; --k) stack.push_back(cells[c].child[k]);
    }
    return out;
  }

  void refine_uniform(int levels) {
    for (int l = 0; l < levels; ++l) {
      auto lv = leaves_dfs();
      for (int c : lv) refine_cell(c);
    }
  }

  // 2:1 balance across faces (Trixi `rebalance!`): refine a leaf when a same-level neighbour's
  // child adjacent to the shared face has children itself.
  void rebalance() {
    bool changed = true;
    while (changed) {
      changed = false;
      auto lv = leaves_dfs();
      for (int c : lv) {
        bool need = false;
        for (int dir = 1; dir <= 2 * ndim && !need; ++dir) {
          bool coarse;
          int nb = neighbor(c, dir, coarse);
          if (nb < 0 || !has_children(nb)) continue;
          int d = (dir - 1) / 2;
          int facing = (dir % 2 == 0) ? 0 : 1;  // neighbour children on the side facing c
          for (int k = 0; k < nchild; ++k)
            if (((k >> d) & 1) == facing && has_children(cells[nb].child[k])) { need = true; break; }
        }
        if (need) { refine_cell(c); changed = true; }
      }
    }
  }

  // Trixi `refine_box!`: refine all leaves whose centre is strictly inside the box, then rebalance
  void refine_box(const RefinementBox& b) {
    auto lv = leaves_dfs();
    for (int c : lv) {
      bool in = true;
      for (int d = 0; d < ndim; ++d)
        in = in && (b.lo[d] < cells[c].x[d]) && (cells[c].x[d] < b.hi[d]);
      if (in) refine_cell(c);
    }
    rebalance();
  }
};

struct Containers {
  int ndim = 3, N = 4;
  // elements
  int64_t nelements = 0;
  std::vector<int64_t> cell_levels;        // [E]
  std::vector<int64_t> cell_icoords;       // [3,E] integer coords at own level (diagnostic)
  vec cell_centers;                        // [ndim,E]
  vec inverse_jacobian;                    // [E]
  vec node_coordinates;                    // [ndim, N^ndim, E]
  // interfaces
  int64_t ninterfaces = 0;
  std::vector<int64_t> if_neighbor_ids;    // [2,I]
  std::vector<int64_t> if_orientations;    // [I]
  // boundaries
  int64_t nboundaries = 0;
  std::vector<int64_t> bd_neighbor_ids, bd_orientations, bd_neighbor_sides;  // [B]
  std::vector<int64_t> n_boundaries_per_direction;                           // [2*ndim]
  vec bd_node_coordinates;                                                   // [ndim, N^(ndim-1), B]
  // mortars
  int64_t nmortars = 0;
  std::vector<int64_t> mo_neighbor_ids;    // [2^(ndim-1)+1, M]
  std::vector<int64_t> mo_large_sides, mo_orientations;  // [M]
};

inline Containers build_containers(const Tree& t, const Basis& b) {
  Containers c;
  const int nd = t.ndim, N = b.N;
  c.ndim = nd; c.N = N;
  std::vector<int> leaves = t.leaves_dfs();
  int64_t E = (int64_t)leaves.size();
  c.nelements = E;
  std::vector<int64_t> c2e(t.cells.size(), 0);
  for (int64_t e = 0; e < E; ++e) c2e[leaves[e]] = e + 1;

  int nn = 1; for (int d = 0; d < nd; ++d) nn *= N;
  int nf = nn / N;
  c.cell_levels.resize(E); c.cell_icoords.assign(3 * E, 0); c.cell_centers.resize((size_t)nd * E);
  c.inverse_jacobian.resize(E);
  c.node_coordinates.resize((size_t)nd * nn * E);
  for (int64_t e = 0; e < E; ++e) {
    const Cell& cl = t.cells[leaves[e]];
    c.cell_levels[e] = cl.level;
    for (int d = 0; d < 3; ++d) c.cell_icoords[3 * e + d] = cl.ic[d];
    double dx = t.length_at_level(cl.level);
    double jac = dx / 2;
    c.inverse_jacobian[e] = 1.0 / jac;
    for (int d = 0; d < nd; ++d) c.cell_centers[nd * e + d] = cl.x[d];
    for (int n = 0; n < nn; ++n) {
      int idx[3] = {n % N, (n / N) % N, n / (N * N)};
      for (int d = 0; d < nd; ++d)
        c.node_coordinates[d + (size_t)nd * (n + (size_t)nn * e)] = cl.x[d] + jac * b.nodes[idx[d]];
    }
  }
  // interfaces: element-outer, positive directions only, neighbour must be a leaf
  for (int64_t e = 0; e < E; ++e)
    for (int dir = 2; dir <= 2 * nd; dir += 2) {
      bool coarse;
      int nb = t.neighbor(leaves[e], dir, coarse);
      if (nb < 0 || t.has_children(nb)) continue;
      c.if_neighbor_ids.push_back(e + 1);
      c.if_neighbor_ids.push_back(c2e[nb]);
      c.if_orientations.push_back(dir / 2);
    }
  c.ninterfaces = (int64_t)c.if_orientations.size();
  // boundaries: direction-outer, element-inner
  c.n_boundaries_per_direction.assign(2 * nd, 0);
  for (int dir = 1; dir <= 2 * nd; ++dir)
    for (int64_t e = 0; e < E; ++e) {
      bool coarse;
      int nb = t.neighbor(leaves[e], dir, coarse);
      if (nb >= 0 || coarse) continue;
      c.bd_neighbor_ids.push_back(e + 1);
      c.bd_neighbor_sides.push_back(dir % 2 == 0 ? 1 : 2);
      c.bd_orientations.push_back((dir + 1) / 2);
      c.n_boundaries_per_direction[dir - 1] += 1;
      // face node coordinates
      int d = (dir - 1) / 2;
      int fixed = (dir % 2 == 0) ? N - 1 : 0;
      for (int f = 0; f < nf; ++f) {
        int a = f % N, bb = f / N;  // face indices over remaining dims in increasing order
        int idx[3] = {0, 0, 0};
        int q = 0;
        for (int dd = 0; dd < nd; ++dd) {
          if (dd == d) idx[dd] = fixed;
          else { idx[dd] = (q == 0) ? a : bb; ++q; }
        }
        int n = idx[0] + N * (idx[1] + N * idx[2]);
        for (int dd = 0; dd < nd; ++dd)
          c.bd_node_coordinates.push_back(c.node_coordinates[dd + (size_t)nd * (n + (size_t)nn * e)]);
      }
    }
  c.nboundaries = (int64_t)c.bd_neighbor_ids.size();
  // mortars: element-outer, all directions, neighbour has children -> this element is the large one
  // small-children tables (1-based child numbers) per direction, see SURVEY.md A.3
  static const int ch3[6][4] = {{2, 4, 6, 8}, {1, 3, 5, 7}, {3, 4, 7, 8}, {1, 2, 5, 6}, {5, 6, 7, 8}, {1, 2, 3, 4}};
  static const int ch2[4][2] = {{2, 4}, {1, 3}, {3, 4}, {1, 2}};
  if (nd >= 2)
    for (int64_t e = 0; e < E; ++e)
      for (int dir = 1; dir <= 2 * nd; ++dir) {
        bool coarse;
        int nb = t.neighbor(leaves[e], dir, coarse);
        if (nb < 0 || !t.has_children(nb)) continue;
        int ns = (nd == 3) ? 4 : 2;
        for (int s = 0; s < ns; ++s) {
          int k = (nd == 3 ? ch3[dir - 1][s] : ch2[dir - 1][s]) - 1;
          int ch = t.cells[nb].child[k];
          if (t.has_children(ch)) throw std::runtime_error("mesh not 2:1 balanced");
          c.mo_neighbor_ids.push_back(c2e[ch]);
        }
        c.mo_neighbor_ids.push_back(e + 1);
        c.mo_large_sides.push_back(dir % 2 == 0 ? 1 : 2);
        c.mo_orientations.push_back((dir + 1) / 2);
      }
  c.nmortars = (int64_t)c.mo_orientations.size();
  return c;
}

}  // namespace orc
