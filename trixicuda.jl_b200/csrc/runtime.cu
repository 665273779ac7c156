// libtrixib200 host runtime: handle construction (device containers, Morton-range partition, halo plans),
// rhs! orchestration on CUDA streams, and the C ABI declared in include/trixib200.h.
// Replaces the reference's L0-L3 (src/solvers/cache.jl, containers_*.jl, dg_*.jl host launchers,
// src/auxiliary/configurators.jl) -- see DESIGN.md.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cstdint>
#include <string>
#include <vector>
#include "../../include/trixib200.h"
#include "device.cuh"
#include "kernels_staged.cuh"
#include "kernels_fused.cuh"
#include "kernels_warp3d.cuh"
#include "kernels_line3d.cuh"
#include "kernels_line6.cuh"
#include "kernels_analysis.cuh"

using namespace tb;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CUDA_TRY(x)                                                                              \
  do {                                                                                           \
    cudaError_t e_ = (x);                                                                        \
    if (e_ != cudaSuccess)                                                                       \
      return fail(TRIXIB200_ECUDA, std::string(#x) + ": " + cudaGetErrorString(e_));             \
  } while (0)

// ---------------------------------------------------------------------------------------------- NCCL (lazy)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
struct Nccl {
  void* lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*CommAbort)(ncclComm_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool load() {
    if (lib) return true;
    // RTLD_NOLOAD first: reuse the libnccl the host process (e.g. PyTorch) already mapped
    lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW);
    if (!lib) return false;
#define SYM(n) *(void**)(&n) = dlsym(lib, "nccl" #n)
    SYM(GetUniqueId); SYM(CommInitRank); SYM(CommDestroy); SYM(CommAbort); SYM(Send); SYM(Recv); SYM(AllReduce); SYM(AllGather);
    SYM(GroupStart); SYM(GroupEnd); SYM(GetErrorString);
#undef SYM
    return GetUniqueId && CommInitRank && Send && Recv && AllReduce && GroupStart && GroupEnd;
  }
};
static Nccl g_nccl;
constexpr int NCCL_FLOAT64 = 8, NCCL_SUM = 0, NCCL_MAX = 2;
constexpr size_t P2P_FLAG_BYTES = 512;    // in-kernel halo exchange: 64 flag words, rank r raises word r in its peers' memory


// ---------------------------------------------------------------------------------------------- partition plan
// Pure host logic (no CUDA): contiguous range of the leaf (Morton) order per rank, local interface list, halo
// send/recv plan and the per-element face neighbour table. Exposed through trixib200_plan_* so the N>1 host
// logic can be tested on CPU-only machines.
struct Plan {
  int nd = 0, nsmall = 0;
  int64_t EG = 0, first = 0, last = 0;
  std::vector<int> if_left, if_right, if_dim;         // local interfaces (halo side = halo code)
  std::vector<int64_t> if_global;                      // global interface index of each local interface
  std::vector<int> face_nbr;                           // [E, 2*nd]
  std::vector<int> send_elem, send_dir;                // peer-major send list (local element, 0-based direction)
  std::vector<int64_t> send_global_iface;              // global interface index of each send entry
  std::vector<int> peers;
  std::vector<int64_t> peer_count;                     // faces RECEIVED from each peer (interface faces, then mortar faces)
  std::vector<int64_t> peer_send_count;                // faces SENT to each peer (equal to peer_count without cut mortars)
  int64_t nrecv = 0;
  std::vector<int> elems_interior, elems_halo;
  std::vector<int> bd_elem, bd_dim, bd_side, bd_dir;
  std::vector<int64_t> bd_global;
  std::vector<int> mo_ids, mo_side, mo_dim;
  std::vector<int64_t> mo_global;                      // global mortar index of each local (possibly replicated) mortar
};

static int64_t range_first(int64_t EG, int nranks, int r) { return (int64_t)((__int128)EG * r / nranks); }

static int build_plan(const trixib200_config& c, const trixib200_mesh_host* ms, Plan& P) {
  const int nd = c.ndim;
  P.nd = nd;
  P.nsmall = nd == 3 ? 4 : (nd == 2 ? 2 : 0);
  const int64_t EG = ms->nelements;
  P.EG = EG;
  if (c.nranks < 1 || c.rank < 0 || c.rank >= c.nranks) return fail(TRIXIB200_EINVAL, "bad rank/nranks");
  if (EG < c.nranks) return fail(TRIXIB200_EINVAL, "fewer elements than ranks");
  const int64_t first = range_first(EG, c.nranks, c.rank), last = range_first(EG, c.nranks, c.rank + 1);
  P.first = first; P.last = last;
  const int64_t E = last - first;
  if (E >= (int64_t)1 << 30) return fail(TRIXIB200_EUNSUPPORTED, "too many elements per rank for int32 ids");
  auto owner = [&](int64_t g) {
    int r = (int)(((__int128)(g + 1) * c.nranks - 1) / EG);
    while (r > 0 && range_first(EG, c.nranks, r) > g) --r;
    while (r + 1 < c.nranks && range_first(EG, c.nranks, r + 1) <= g) ++r;
    return r;
  };
  P.face_nbr.assign((size_t)E * 2 * nd, NB_SFV);
  std::vector<std::vector<int>> send_elem(c.nranks), send_dir(c.nranks);
  std::vector<std::vector<int64_t>> send_gi(c.nranks);
  struct HaloRef { int64_t iface; int peer; int64_t k; };
  std::vector<HaloRef> halos;
  std::vector<int64_t> per_peer(c.nranks, 0);
  for (int64_t s = 0; s < ms->ninterfaces; ++s) {
    int64_t L = ms->interfaces_neighbor_ids[2 * s] - 1, R = ms->interfaces_neighbor_ids[2 * s + 1] - 1;
    int dim = (int)ms->interfaces_orientations[s] - 1;
    if (L < 0 || L >= EG || R < 0 || R >= EG || dim < 0 || dim >= nd) return fail(TRIXIB200_EINVAL, "bad interface entry");
    bool lin = L >= first && L < last, rin = R >= first && R < last;
    if (!lin && !rin) continue;
    P.if_global.push_back(s);
    if (lin && rin) {
      P.if_left.push_back((int)(L - first)); P.if_right.push_back((int)(R - first)); P.if_dim.push_back(dim);
      P.face_nbr[(size_t)(L - first) * 2 * nd + 2 * dim + 1] = (int)(R - first);
      P.face_nbr[(size_t)(R - first) * 2 * nd + 2 * dim] = (int)(L - first);
    } else {
      int peer = owner(lin ? R : L);
      int64_t k = per_peer[peer]++;
      halos.push_back({(int64_t)P.if_left.size(), peer, k});
      send_gi[peer].push_back(s);
      if (lin) {
        P.if_left.push_back((int)(L - first)); P.if_right.push_back(0); P.if_dim.push_back(dim);
        send_elem[peer].push_back((int)(L - first)); send_dir[peer].push_back(2 * dim + 1);
      } else {
        P.if_left.push_back(0); P.if_right.push_back((int)(R - first)); P.if_dim.push_back(dim);
        send_elem[peer].push_back((int)(R - first)); send_dir[peer].push_back(2 * dim);
      }
    }
  }
  // ---- mortars whose elements live on several ranks are REPLICATED on every rank that owns one of them (SURVEY.md
  // section 8(e)): the rank receives the face traces of the elements it does not own -- the same unit as a cut
  // interface face -- and computes the whole mortar; it only keeps the fluxes of its own elements. Sender and
  // receiver walk the global mortar list in the same order, so the k-th mortar face exchanged between two ranks has
  // the same k on both sides; per peer the mortar faces come behind the interface faces.
  const int rows = P.nsmall + 1;
  struct MortarRecv { int64_t mortar_local; int row; int peer; int64_t k; };
  std::vector<MortarRecv> mrecv;
  std::vector<int64_t> recv_per_peer(per_peer);         // interface faces: one received per one sent
  std::vector<char> touches_halo(E, 0);
  for (int64_t m = 0; m < ms->nmortars; ++m) {
    int own[5];
    int64_t gid[5];
    bool mine = false;
    for (int r = 0; r < rows; ++r) {
      gid[r] = ms->mortars_neighbor_ids[rows * m + r] - 1;
      if (gid[r] < 0 || gid[r] >= EG) return fail(TRIXIB200_EINVAL, "bad mortar entry");
      own[r] = (gid[r] >= first && gid[r] < last) ? c.rank : owner(gid[r]);
      mine = mine || own[r] == c.rank;
    }
    if (!mine) continue;
    const int o1 = (int)ms->mortars_orientations[m], ls = (int)ms->mortars_large_sides[m];
    if (o1 < 1 || o1 > nd || (ls != 1 && ls != 2)) return fail(TRIXIB200_EINVAL, "bad mortar entry");
    const int dir_small = 2 * o1 + ls - 3, dir_large = 2 * o1 - ls;      // 0-based face directions towards the mortar
    std::vector<int> involved(own, own + rows);
    std::sort(involved.begin(), involved.end());
    involved.erase(std::unique(involved.begin(), involved.end()), involved.end());
    const int64_t ml = (int64_t)P.mo_side.size();
    for (int r = 0; r < rows; ++r) {
      if (own[r] == c.rank) {
        const int le = (int)(gid[r] - first);
        P.mo_ids.push_back(le);
        for (int b : involved)
          if (b != c.rank) {
            send_elem[b].push_back(le); send_dir[b].push_back(r == rows - 1 ? dir_large : dir_small);
            send_gi[b].push_back(-1);
            touches_halo[le] = 1;
          }
      } else {
        P.mo_ids.push_back(0);                                           // halo code filled in below
        mrecv.push_back({ml, r, own[r], recv_per_peer[own[r]]++});
      }
    }
    P.mo_side.push_back(ls);
    P.mo_dim.push_back(o1 - 1);
    P.mo_global.push_back(m);
  }
  // halo slots are peer-major (receive order); the k-th interface face exchanged with a peer has the same k on both
  // sides because both ranks walk the global interface list in the same order
  std::vector<int64_t> peer_off(c.nranks + 1, 0);
  for (int p = 0; p < c.nranks; ++p) peer_off[p + 1] = peer_off[p] + recv_per_peer[p];
  P.nrecv = peer_off[c.nranks];
  for (int p = 0; p < c.nranks; ++p)
    if (recv_per_peer[p] > 0 || !send_elem[p].empty()) {
      P.peers.push_back(p); P.peer_count.push_back(recv_per_peer[p]); P.peer_send_count.push_back((int64_t)send_elem[p].size());
    }
  for (int p = 0; p < c.nranks; ++p) {
    P.send_elem.insert(P.send_elem.end(), send_elem[p].begin(), send_elem[p].end());
    P.send_dir.insert(P.send_dir.end(), send_dir[p].begin(), send_dir[p].end());
    P.send_global_iface.insert(P.send_global_iface.end(), send_gi[p].begin(), send_gi[p].end());
  }
  for (const MortarRecv& mr : mrecv) {
    int64_t slot = peer_off[mr.peer] + mr.k;
    if (slot >= ((int64_t)1 << 30)) return fail(TRIXIB200_EUNSUPPORTED, "too many halo faces");
    P.mo_ids[(size_t)rows * mr.mortar_local + mr.row] = nb_from_halo_slot((int)slot);
  }
  for (const HaloRef& hr : halos) {
    int64_t slot = peer_off[hr.peer] + hr.k;
    if (slot >= ((int64_t)1 << 30)) return fail(TRIXIB200_EUNSUPPORTED, "too many halo faces");
    int code = nb_from_halo_slot((int)slot);
    int64_t s = hr.iface;
    int dim = P.if_dim[s];
    int le = send_elem[hr.peer][hr.k], ldir = send_dir[hr.peer][hr.k];
    if (ldir == 2 * dim + 1) P.if_right[s] = code; else P.if_left[s] = code;
    P.face_nbr[(size_t)le * 2 * nd + ldir] = code;
    touches_halo[le] = 1;
  }
  for (int64_t e = 0; e < E; ++e) (touches_halo[e] ? P.elems_halo : P.elems_interior).push_back((int)e);
  // boundaries (sorted by direction; keep only local ones, order preserved)
  int64_t b = 0;
  for (int dir = 0; dir < 2 * nd; ++dir) {
    int64_t nb = ms->nboundaries > 0 ? ms->n_boundaries_per_direction[dir] : 0;
    if (nb > 0 && c.boundary_conditions[dir] == TRIXIB200_BC_PERIODIC)
      return fail(TRIXIB200_EINVAL, "mesh has a non-periodic boundary in a direction whose boundary condition is periodic");
    for (int64_t k = 0; k < nb; ++k, ++b) {
      int64_t g = ms->boundaries_neighbor_ids[b] - 1;
      if (g < first || g >= last) continue;
      P.bd_elem.push_back((int)(g - first));
      P.bd_dim.push_back((int)ms->boundaries_orientations[b] - 1);
      P.bd_side.push_back((int)ms->boundaries_neighbor_sides[b]);
      P.bd_dir.push_back(dir);
      P.bd_global.push_back(b);
    }
  }
  if (b != ms->nboundaries) return fail(TRIXIB200_EINVAL, "n_boundaries_per_direction does not sum to nboundaries");
  return 0;
}

extern "C" int trixib200_plan_create(const trixib200_config* cfg, const trixib200_mesh_host* mesh, void** out) {
  if (!cfg || !mesh || !out) return fail(TRIXIB200_EINVAL, "null argument");
  Plan* P = new Plan();
  int rc = build_plan(*cfg, mesh, *P);
  if (rc) { delete P; *out = nullptr; return rc; }
  *out = P;
  return 0;
}
extern "C" int trixib200_plan_destroy(void* p) { delete (Plan*)p; return 0; }
static bool plan_array(Plan* P, const std::string& n, const void** data, int64_t* len, int* width) {
#define PA(name, vec, w) if (n == name) { *data = P->vec.data(); *len = (int64_t)P->vec.size(); *width = w; return true; }
  PA("if_left", if_left, 4) PA("if_right", if_right, 4) PA("if_dim", if_dim, 4) PA("if_global", if_global, 8)
  PA("face_nbr", face_nbr, 4) PA("send_elem", send_elem, 4) PA("send_dir", send_dir, 4)
  PA("send_global_iface", send_global_iface, 8) PA("peers", peers, 4) PA("peer_count", peer_count, 8)
  PA("peer_send_count", peer_send_count, 8) PA("mo_side", mo_side, 4) PA("mo_dim", mo_dim, 4) PA("mo_global", mo_global, 8)
  PA("elems_interior", elems_interior, 4) PA("elems_halo", elems_halo, 4) PA("bd_elem", bd_elem, 4)
  PA("bd_global", bd_global, 8) PA("mo_ids", mo_ids, 4)
#undef PA
  return false;
}
extern "C" int64_t trixib200_plan_len(void* p, const char* name) {
  const void* dptr; int64_t len; int w;
  if (!p || !name) return -1;
  std::string n(name);
  if (n == "first_element") return ((Plan*)p)->first;
  if (n == "nelements") return ((Plan*)p)->last - ((Plan*)p)->first;
  if (!plan_array((Plan*)p, n, &dptr, &len, &w)) return -1;
  return len;
}
// copies the named plan array widened to int64
extern "C" int trixib200_plan_get(void* p, const char* name, int64_t* out, int64_t n) {
  const void* dptr; int64_t len; int w;
  if (!p || !name || !plan_array((Plan*)p, name, &dptr, &len, &w) || len != n) return fail(TRIXIB200_EINVAL, "bad plan array request");
  for (int64_t i = 0; i < n; ++i) out[i] = (w == 4) ? (int64_t)((const int*)dptr)[i] : ((const int64_t*)dptr)[i];
  return 0;
}

// ---------------------------------------------------------------------------------------------- handle
struct trixib200_handle {
  trixib200_config cfg;
  Dev d;
  int nsmall = 0;
  int64_t E_global = 0, first = 0;
  cudaStream_t stream = nullptr, comm_stream = nullptr;
  bool owns_stream = false;
  cudaEvent_t ev_pack = nullptr, ev_halo = nullptr, ev_a = nullptr, ev_b = nullptr;
  std::vector<void*> allocs;
  double* d_scalar = nullptr;
  int64_t launches = 0;
  bool fused = false, warp3d = false, line3d = false;
  LineOps line_ops;
  // fused-path element lists (multi-GPU overlap): interior first, then elements touching a halo face
  int* d_elems_interior = nullptr; int* d_elems_halo = nullptr;
  int64_t n_interior = 0, n_halo_elems = 0;
  // halo plan
  std::vector<int> peers;                 // peer ranks
  std::vector<int64_t> peer_count;        // faces received from each peer
  std::vector<int64_t> peer_send_count;   // faces sent to each peer (== peer_count unless mortars cross a cut)
  ncclComm_t comm = nullptr;
  int sm_count = 148;
  // in-kernel halo exchange over peer memory (CUDA IPC; see P2PArgs in kernels_line6.cuh)
  struct P2P {
    bool enabled = false;
    bool mesh_ok = false;                      // no boundary / mortar faces anywhere in the GLOBAL mesh (same on all ranks)
    std::string why = "not set up";            // why it is off (reported by trixib200_last_error after size("p2p_halo"))
    unsigned char* base = nullptr;             // local: [flags P2P_FLAG_BYTES][recv buffer epoch-parity 0][parity 1]
    size_t buf_doubles = 0;
    unsigned long long epoch = 0;
    std::vector<void*> peer_base;              // opened IPC mappings of the peers' allocations
    int* d_peer_first = nullptr; int* d_peer_rank = nullptr;
    double** d_peer_dst[2] = {nullptr, nullptr};
    unsigned long long** d_peer_flag = nullptr;
    unsigned int* d_done = nullptr;
    int* d_elems_all = nullptr;                // interior elements, then the elements that touch a cut face
    std::vector<int64_t> slot_off;             // [nranks] first halo slot of every peer rank (-1: not a peer)
  } p2p;
  // rhs_host: library-owned device mirrors of the caller's host vectors, and the chunk pipeline
  double* host_u = nullptr; double* host_du = nullptr;
  double* an_buf = nullptr;                // analysis kernels: operators + per-CTA partials (lazy)
  double* du_scratch = nullptr;            // rk2n_stage on kernel families without the fused epilogue (lazy)
  std::vector<int> face_nbr_host;          // kept for the chunk dependency analysis
  std::vector<double> chunk_key_src;       // last-dimension coordinate of every element (upload ordering)
  struct HostPipe {
    bool built = false, usable = false;
    int64_t ce = 0, nchunks = 0;
    struct Run { int64_t first, count; };   // contiguous elements of the Morton order
    std::vector<std::vector<Run>> runs;    // per slab
    std::vector<int64_t> slab_off;         // [nslabs + 1] offsets into d_iota (elements grouped by slab)
    std::vector<std::vector<int>> ready;   // ready[i]: slabs computable once slab i has landed
    std::vector<cudaEvent_t> ev_in, ev_done;
    cudaStream_t s_in = nullptr, s_out = nullptr;
    int* d_iota = nullptr;                 // element ids grouped by slab
  } pipe;
};

template <class T> static int upload(trixib200_handle* h, const std::vector<T>& v, T** out) {
  *out = nullptr;
  size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) return fail(TRIXIB200_ENOMEM, "cudaMalloc failed (" + std::to_string(bytes) + " B)");
  h->allocs.push_back(p);
  if (!v.empty()) CUDA_TRY(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  *out = (T*)p;
  return 0;
}
static int dalloc(trixib200_handle* h, size_t n, double** out, bool zero = true) {
  void* p = nullptr;
  size_t bytes = std::max<size_t>(n, 1) * sizeof(double);
  if (cudaMalloc(&p, bytes) != cudaSuccess) return fail(TRIXIB200_ENOMEM, "cudaMalloc failed (" + std::to_string(bytes) + " B)");
  h->allocs.push_back(p);
  if (zero) CUDA_TRY(cudaMemset(p, 0, bytes));
  *out = (double*)p;
  return 0;
}

static inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

// dispatch on (equations, ndim)
#define TB_DISPATCH_EQ(h, ...)                                                          \
  do {                                                                                  \
    const int nd_ = (h)->cfg.ndim;                                                      \
    switch ((h)->cfg.equations) {                                                       \
      case TRIXIB200_EQ_ADVECTION:                                                      \
        if (nd_ == 1) { using Eq = EqAdvection<1>; __VA_ARGS__; }                              \
        else if (nd_ == 2) { using Eq = EqAdvection<2>; __VA_ARGS__; }                         \
        else { using Eq = EqAdvection<3>; __VA_ARGS__; }                                       \
        break;                                                                          \
      case TRIXIB200_EQ_EULER:                                                          \
        if (nd_ == 1) { using Eq = EqEuler<1>; __VA_ARGS__; }                                  \
        else if (nd_ == 2) { using Eq = EqEuler<2>; __VA_ARGS__; }                             \
        else { using Eq = EqEuler<3>; __VA_ARGS__; }                                           \
        break;                                                                          \
      default: { using Eq = EqMhd3; __VA_ARGS__; } break;                                      \
    }                                                                                   \
  } while (0)
#define TB_DISPATCH_ND(h, ...)                                                          \
  do {                                                                                  \
    const int nd_ = (h)->cfg.ndim;                                                      \
    if (nd_ == 1) { constexpr int ND = 1; __VA_ARGS__; }                                       \
    else if (nd_ == 2) { constexpr int ND = 2; __VA_ARGS__; }                                  \
    else { constexpr int ND = 3; __VA_ARGS__; }                                                \
  } while (0)

// enumerated initial conditions per equation set (TRIXIB200_IC_NONE = -1: the caller's IC is not enumerated; the
// device then refuses Dirichlet(ic) boundaries, fill_initial_condition and calc_error_norms instead of silently
// substituting another state)
static bool ic_supported(const trixib200_config& c) {
  const int ic = c.initial_condition;
  if (ic < TRIXIB200_IC_CONSTANT || ic > TRIXIB200_IC_DENSITY_WAVE) return false;
  if (ic == TRIXIB200_IC_DENSITY_WAVE) return c.equations == TRIXIB200_EQ_EULER;
  if (ic == TRIXIB200_IC_WEAK_BLAST_WAVE) return c.equations != TRIXIB200_EQ_ADVECTION;
  return true;
}
static bool flux_supported(const trixib200_config& c, int k) {
  if (c.equations == TRIXIB200_EQ_ADVECTION) return EqAdvection<1>::supports_flux(k);
  if (c.equations == TRIXIB200_EQ_EULER) return EqEuler<1>::supports_flux(k);
  return EqMhd3::supports_flux(k);
}
static bool flux_symmetric(int k) {
  return k == TRIXIB200_FLUX_CENTRAL || k == TRIXIB200_FLUX_RANOCHA || k == TRIXIB200_FLUX_SHIMA_ETAL ||
         k == TRIXIB200_FLUX_HINDENLANG_GASSNER;
}

// ---------------------------------------------------------------------------------------------- create
extern "C" const char* trixib200_last_error(void) { return g_err.c_str(); }
extern "C" int trixib200_version(void) { return 100; }

extern "C" int trixib200_destroy(trixib200_handle* h) {
  if (!h) return 0;
  cudaSetDevice(h->cfg.device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  // Teardown must not depend on what the other ranks are doing: handles are destroyed by garbage collectors and at
  // interpreter exit, in no agreed order (an 8-rank run that printed its result hung in here until it was killed,
  // profiles/r2_pytest_multi_8.log). Hence: the stream is idle (synchronised above), so the communicator is released
  // with ncclCommAbort (local, never waits for a peer; ncclCommDestroy finalises collectively), the peers' buffers are
  // unmapped, and this rank's exported receive buffer is NOT freed while peers may still have it mapped (freeing
  // exported memory before every importer has closed it is undefined): it goes back with the context.
  bool had_peers = false;
  for (void* pb : h->p2p.peer_base) if (pb) { cudaIpcCloseMemHandle(pb); had_peers = true; }
  if (h->p2p.base && !had_peers) cudaFree(h->p2p.base);
  if (h->comm) {
    if (h->cfg.nranks > 1 && g_nccl.CommAbort) g_nccl.CommAbort(h->comm);
    else if (g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  }
  for (void* p : h->allocs) cudaFree(p);
  if (h->ev_pack) cudaEventDestroy(h->ev_pack);
  if (h->ev_halo) cudaEventDestroy(h->ev_halo);
  if (h->ev_a) cudaEventDestroy(h->ev_a);
  if (h->ev_b) cudaEventDestroy(h->ev_b);
  for (cudaEvent_t e : h->pipe.ev_in) cudaEventDestroy(e);
  for (cudaEvent_t e : h->pipe.ev_done) cudaEventDestroy(e);
  if (h->pipe.s_in) cudaStreamDestroy(h->pipe.s_in);
  if (h->pipe.s_out) cudaStreamDestroy(h->pipe.s_out);
  if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
  if (h->stream && h->owns_stream) cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

static int create_impl(const trixib200_config* cfg, const trixib200_basis_host* bs, const trixib200_mesh_host* ms,
                       trixib200_handle* h) {
  const trixib200_config& c = *cfg;
  // ---- validation: clear errors for anything un-enumerated, never a fallback
  if (c.ndim < 1 || c.ndim > 3) return fail(TRIXIB200_EINVAL, "ndim must be 1, 2 or 3");
  if (c.equations < 0 || c.equations > 2) return fail(TRIXIB200_EUNSUPPORTED, "unknown equations id");
  if (c.equations == TRIXIB200_EQ_MHD && c.ndim != 3) return fail(TRIXIB200_EUNSUPPORTED, "IdealGlmMhdEquations only in 3D");
  if (bs->nnodes != c.polydeg + 1) return fail(TRIXIB200_EINVAL, "basis.nnodes != polydeg + 1");
  if (bs->nnodes < 2 || bs->nnodes > MAXN) return fail(TRIXIB200_EUNSUPPORTED, "polydeg must be in 1..7");
  if (c.volume_integral < 0 || c.volume_integral > 2) return fail(TRIXIB200_EUNSUPPORTED, "unknown volume integral");
  if (!flux_supported(c, c.surface_flux)) return fail(TRIXIB200_EUNSUPPORTED, "surface_flux not available for these equations");
  if (c.volume_integral != TRIXIB200_VI_WEAK_FORM) {
    if (!flux_supported(c, c.volume_flux)) return fail(TRIXIB200_EUNSUPPORTED, "volume_flux not available for these equations");
    if (!flux_symmetric(c.volume_flux))
      return fail(TRIXIB200_EUNSUPPORTED, "flux differencing needs a symmetric two-point volume flux");
  }
  if (c.volume_integral == TRIXIB200_VI_SHOCK_CAPTURING_HG) {
    if (!flux_supported(c, c.volume_flux_fv)) return fail(TRIXIB200_EUNSUPPORTED, "volume_flux_fv not available for these equations");
    if (c.equations == TRIXIB200_EQ_ADVECTION) return fail(TRIXIB200_EUNSUPPORTED, "shock capturing needs Euler or MHD");
  }
  if (c.nonconservative && c.equations != TRIXIB200_EQ_MHD)
    return fail(TRIXIB200_EUNSUPPORTED, "nonconservative terms only for IdealGlmMhdEquations3D (flux_nonconservative_powell)");
  if (c.equations == TRIXIB200_EQ_MHD && !c.nonconservative)
    return fail(TRIXIB200_EUNSUPPORTED, "IdealGlmMhdEquations3D requires (flux, flux_nonconservative_powell) tuples");
  if (c.source_terms != TRIXIB200_SRC_NONE && c.equations != TRIXIB200_EQ_EULER)
    return fail(TRIXIB200_EUNSUPPORTED, "source_terms_convergence_test only for compressible Euler");
  if (c.nranks < 1 || c.rank < 0 || c.rank >= c.nranks) return fail(TRIXIB200_EINVAL, "bad rank/nranks");
  if (ms->nelements < c.nranks) return fail(TRIXIB200_EINVAL, "fewer elements than ranks");
  for (int q = 0; q < 2 * c.ndim; ++q) {
    const int bc = c.boundary_conditions[q];
    if (bc < TRIXIB200_BC_PERIODIC || bc > TRIXIB200_BC_SLIP_WALL) return fail(TRIXIB200_EUNSUPPORTED, "unknown boundary condition id");
    if (bc == TRIXIB200_BC_DIRICHLET_IC && !ic_supported(c))
      return fail(TRIXIB200_EUNSUPPORTED, "BoundaryConditionDirichlet needs an enumerated initial condition "
                                          "(config.initial_condition is TRIXIB200_IC_NONE or unknown for these equations)");
    if (bc == TRIXIB200_BC_SLIP_WALL && c.equations != TRIXIB200_EQ_EULER)
      return fail(TRIXIB200_EUNSUPPORTED, "boundary_condition_slip_wall only for compressible Euler");
  }

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(TRIXIB200_ECUDA, "no CUDA device available (libtrixib200 has no CPU fallback)");
  if (c.device < 0 || c.device >= ndev) return fail(TRIXIB200_EINVAL, "bad device ordinal");
  CUDA_TRY(cudaSetDevice(c.device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, c.device));
  h->sm_count = prop.multiProcessorCount;
  h->cfg = c;
  CUDA_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  h->owns_stream = true;
  CUDA_TRY(cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreateWithFlags(&h->ev_pack, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&h->ev_halo, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreate(&h->ev_a));
  CUDA_TRY(cudaEventCreate(&h->ev_b));

  const int nd = c.ndim, N = bs->nnodes;
  Dev& d = h->d;
  std::memset(&d, 0, sizeof(d));
  d.ndim = nd; d.N = N;
  d.nn = 1; for (int q = 0; q < nd; ++q) d.nn *= N;
  d.nf = d.nn / N;
  d.nv = c.equations == TRIXIB200_EQ_ADVECTION ? 1 : (c.equations == TRIXIB200_EQ_EULER ? nd + 2 : 9);
  h->nsmall = nd == 3 ? 4 : (nd == 2 ? 2 : 0);

  // ---- operators
  Ops ops;
  std::memset(&ops, 0, sizeof(ops));
  for (int i = 0; i < N; ++i) { ops.nodes[i] = bs->nodes[i]; ops.weights[i] = bs->weights[i]; ops.inv_w[i] = bs->inverse_weights[i]; }
  auto cpm = [&](double* dst, const double* src) { if (src) for (int i = 0; i < N * N; ++i) dst[i] = src[i]; };
  cpm(ops.Dhat, bs->derivative_dhat); cpm(ops.Dsplit, bs->derivative_split); cpm(ops.invV, bs->inverse_vandermonde_legendre);
  cpm(ops.fwd_u, bs->forward_upper); cpm(ops.fwd_l, bs->forward_lower); cpm(ops.rev_u, bs->reverse_upper); cpm(ops.rev_l, bs->reverse_lower);
  ops.factor_1 = bs->boundary_interpolation[0];
  ops.factor_2 = bs->boundary_interpolation[(N - 1) + N * 1];
  if (ms->nmortars > 0 && (!bs->forward_upper || !bs->forward_lower || !bs->reverse_upper || !bs->reverse_lower))
    return fail(TRIXIB200_EINVAL, "mesh has mortars but mortar operators are NULL");
  if (c.volume_integral == TRIXIB200_VI_SHOCK_CAPTURING_HG && !bs->inverse_vandermonde_legendre)
    return fail(TRIXIB200_EINVAL, "shock capturing needs inverse_vandermonde_legendre");
  {
    std::vector<Ops> v(1, ops);
    Ops* p;
    if (int rc = upload(h, v, &p)) return rc;
    d.ops = p;
  }

  // ---- partition plan (host) and its upload
  Plan P;
  if (int rc = build_plan(c, ms, P)) return rc;
  const int64_t EG = P.EG, first = P.first, last = P.last, E = last - first;
  h->E_global = EG; h->first = first;
  d.E = E;
  if ((int64_t)d.nn * E * d.nv >= (int64_t)1 << 40) return fail(TRIXIB200_EUNSUPPORTED, "partition too large");
  h->peers = P.peers; h->peer_count = P.peer_count; h->peer_send_count = P.peer_send_count;

  // ---- elements
  {
    std::vector<double> ij(ms->inverse_jacobian + first, ms->inverse_jacobian + last);
    double* p;
    if (int rc = upload(h, ij, &p)) return rc;
    d.inv_jac = p;
    if (ms->node_coordinates) {
      size_t per = (size_t)nd * d.nn;
      std::vector<double> nc(ms->node_coordinates + per * first, ms->node_coordinates + per * last);
      if (int rc = upload(h, nc, &p)) return rc;
      d.node_coords = p;
    }
    if (ms->cell_centers) {
      std::vector<double> cc(ms->cell_centers + (size_t)nd * first, ms->cell_centers + (size_t)nd * last);
      if (int rc = upload(h, cc, &p)) return rc;
      d.centers = p;
    }
    if (!ms->node_coordinates && !ms->cell_centers && (c.source_terms != TRIXIB200_SRC_NONE))
      return fail(TRIXIB200_EINVAL, "source terms need node_coordinates or cell_centers");
  }
  // ---- interfaces + halo
  d.I = (int64_t)P.if_left.size();
  d.nhalo_send = (int64_t)P.send_elem.size();
  d.nhalo_recv = P.nrecv;
  {
    int* p;
    if (int rc = upload(h, P.if_left, &p)) return rc; d.if_left = p;
    if (int rc = upload(h, P.if_right, &p)) return rc; d.if_right = p;
    if (int rc = upload(h, P.if_dim, &p)) return rc; d.if_dim = p;
    if (int rc = upload(h, P.send_elem, &p)) return rc; d.send_elem = p;
    if (int rc = upload(h, P.send_dir, &p)) return rc; d.send_dir = p;
    double* q;
    if (int rc = dalloc(h, (size_t)d.nv * d.nf * d.nhalo_send, &q)) return rc; d.halo_send = q;
    if (int rc = dalloc(h, (size_t)d.nv * d.nf * d.nhalo_recv, &q)) return rc; d.halo_recv = q;
    // one value per exchanged face: the owner's indicator value, for smoothing across a cut
    if (int rc = dalloc(h, (size_t)d.nhalo_send, &q)) return rc; d.halo_alpha_send = q;
    if (int rc = dalloc(h, (size_t)d.nhalo_recv, &q)) return rc; d.halo_alpha_recv = q;
  }
  // ---- boundaries
  {
    std::vector<double> bco;
    size_t per = (size_t)nd * d.nf;
    for (int64_t gb : P.bd_global) {
      if (!ms->boundaries_node_coordinates) return fail(TRIXIB200_EINVAL, "boundaries need node_coordinates");
      bco.insert(bco.end(), ms->boundaries_node_coordinates + per * gb, ms->boundaries_node_coordinates + per * (gb + 1));
    }
    d.B = (int64_t)P.bd_elem.size();
    int* p; double* q;
    if (int rc = upload(h, P.bd_elem, &p)) return rc; d.bd_elem = p;
    if (int rc = upload(h, P.bd_dim, &p)) return rc; d.bd_dim = p;
    if (int rc = upload(h, P.bd_side, &p)) return rc; d.bd_side = p;
    if (int rc = upload(h, P.bd_dir, &p)) return rc; d.bd_dir = p;
    if (int rc = upload(h, bco, &q)) return rc; d.bd_coords = q;
    if (int rc = dalloc(h, (size_t)2 * d.nv * d.nf * d.B, &q)) return rc; d.boundaries_u = q;
  }
  // ---- mortars
  {
    d.M = (int64_t)P.mo_side.size();
    int* p; double* q;
    if (int rc = upload(h, P.mo_ids, &p)) return rc; d.mo_ids = p;
    if (int rc = upload(h, P.mo_side, &p)) return rc; d.mo_side = p;
    if (int rc = upload(h, P.mo_dim, &p)) return rc; d.mo_dim = p;
    for (int k = 0; k < h->nsmall; ++k) {
      if (int rc = dalloc(h, (size_t)2 * d.nv * d.nf * d.M, &q)) return rc; d.mortar_u[k] = q;
      if (int rc = dalloc(h, (size_t)d.nv * d.nf * d.M, &q)) return rc; d.fstar_p[k] = q;
      if (int rc = dalloc(h, (size_t)d.nv * d.nf * d.M, &q)) return rc; d.fstar_s[k] = q;
    }
  }

  // ---- physics parameters
  d.prm.gamma = c.gamma; d.prm.c_h = c.c_h; d.prm.inv_gm1 = 1.0 / (c.gamma - 1.0);
  for (int q = 0; q < 3; ++q) d.prm.a[q] = c.advection_velocity[q];
  d.volume_integral = c.volume_integral; d.vol_flux = c.volume_flux; d.fv_flux = c.volume_flux_fv;
  d.surf_flux = c.surface_flux; d.noncons = c.nonconservative; d.ic = c.initial_condition; d.src = c.source_terms;
  d.ind_var = c.indicator_variable; d.alpha_smooth = c.alpha_smooth;
  for (int q = 0; q < 6; ++q) d.bc[q] = c.boundary_conditions[q];
  d.alpha_max = c.alpha_max; d.alpha_min = c.alpha_min;

  // ---- fused path availability (3D/2D, polydeg 3) and its element lists
  h->fused = !(c.flags & TRIXIB200_FLAG_STAGED_ONLY) && fused_available(c);
  h->warp3d = h->fused && !(c.flags & TRIXIB200_FLAG_NO_WARP_KERNEL) && warp3d_available(c);
  h->line3d = h->warp3d && !(c.flags & TRIXIB200_FLAG_NO_LINE_KERNEL) && line3d_available(c);
  for (int i = 0; i < 16; ++i) h->line_ops.ds[i] = ops.Dsplit[(i & 3) + N * (i >> 2)];
  h->line_ops.factor_1 = ops.factor_1; h->line_ops.factor_2 = ops.factor_2;
  {
    int* p;
    if (int rc = upload(h, P.face_nbr, &p)) return rc;
    d.face_nbr = p;
    h->face_nbr_host = P.face_nbr;
    h->chunk_key_src.assign((size_t)E, 0.0);
    for (int64_t e = 0; e < E; ++e) {
      if (ms->cell_centers) h->chunk_key_src[e] = ms->cell_centers[(size_t)nd * (first + e) + (nd - 1)];
      else if (ms->node_coordinates) h->chunk_key_src[e] = ms->node_coordinates[(size_t)nd * d.nn * (first + e) + (nd - 1)];
    }
    h->n_interior = (int64_t)P.elems_interior.size(); h->n_halo_elems = (int64_t)P.elems_halo.size();
    if (int rc = upload(h, P.elems_interior, &p)) return rc; h->d_elems_interior = p;
    if (int rc = upload(h, P.elems_halo, &p)) return rc; h->d_elems_halo = p;
    std::vector<int> all(P.elems_interior);
    all.insert(all.end(), P.elems_halo.begin(), P.elems_halo.end());
    if (int rc = upload(h, all, &p)) return rc; h->p2p.d_elems_all = p;
    h->p2p.mesh_ok = ms->nboundaries == 0 && ms->nmortars == 0;
    h->p2p.slot_off.assign(c.nranks, -1);
    int64_t off = 0;
    for (size_t k = 0; k < P.peers.size(); ++k) { h->p2p.slot_off[P.peers[k]] = off; off += P.peer_count[k]; }
  }

  // ---- materialised containers: the staged path needs all of them; the fused path only what boundary /
  // mortar faces write (surface_flux_values) -- and nothing at all on a conforming periodic mesh
  bool need_staged_buffers = !h->fused;
  bool need_sfv = need_staged_buffers || d.B > 0 || d.M > 0;
  {
    double* q;
    if (need_staged_buffers) { if (int rc = dalloc(h, (size_t)2 * d.nv * d.nf * d.I, &q)) return rc; d.interfaces_u = q; }
    if (need_sfv) { if (int rc = dalloc(h, (size_t)d.nv * d.nf * 2 * nd * E, &q)) return rc; d.sfv = q; }
    if (int rc = dalloc(h, (size_t)E, &q)) return rc; d.alpha = q;
    if (int rc = dalloc(h, (size_t)E, &q)) return rc; d.alpha_tmp = q;
    if (int rc = dalloc(h, 8, &q)) return rc; h->d_scalar = q;
  }
  CUDA_TRY(cudaDeviceSynchronize());
  return 0;
}

extern "C" int trixib200_create(const trixib200_config* cfg, const trixib200_basis_host* basis,
                                const trixib200_mesh_host* mesh, trixib200_handle** out) {
  if (!cfg || !basis || !mesh || !out) return fail(TRIXIB200_EINVAL, "null argument");
  *out = nullptr;
  trixib200_handle* h = new trixib200_handle();
  h->cfg = *cfg;
  int rc = create_impl(cfg, basis, mesh, h);
  if (rc != 0) { std::string keep = g_err; trixib200_destroy(h); g_err = keep; return rc; }
  *out = h;
  return 0;
}

extern "C" int64_t trixib200_size(const trixib200_handle* h, const char* name) {
  if (!h || !name) return -1;
  std::string n(name);
  const Dev& d = h->d;
  if (n == "nelements") return d.E;
  if (n == "p2p_halo") { g_err = h->p2p.why; return h->p2p.enabled ? 1 : 0; }
  if (n == "nelements_global") return h->E_global;
  if (n == "first_element") return h->first;
  if (n == "nvars") return d.nv;
  if (n == "nnodes") return d.N;
  if (n == "ndofs") return d.E * d.nn;
  if (n == "nunknowns") return d.E * d.nn * d.nv;
  if (n == "ninterfaces") return d.I;
  if (n == "nboundaries") return d.B;
  if (n == "nmortars") return d.M;
  if (n == "nhalo_faces") return d.nhalo_recv;
  if (n == "fused") return h->fused ? 1 : 0;
  if (n == "warp3d") return h->warp3d ? 1 : 0;
  if (n == "line3d") return h->line3d ? 1 : 0;
  if (n == "npeers") return (int64_t)h->peers.size();
  return -1;
}

// ---------------------------------------------------------------------------------------------- stages
#define LAUNCH(h, kern, n, threads, smem, ...)                                       \
  do {                                                                               \
    if ((n) > 0) {                                                                   \
      kern<<<nblk((n), (threads)), (threads), (smem), (h)->stream>>>(__VA_ARGS__);   \
      (h)->launches++;                                                               \
    }                                                                                \
  } while (0)

static int st_indicator(trixib200_handle* h, const double* u) {
  Dev& d = h->d;
  int threads = ((d.nn + 31) / 32) * 32;
  TB_DISPATCH_EQ(h, { if (d.E > 0) { k_indicator<Eq><<<(unsigned)d.E, threads, 2 * d.nn * sizeof(double), h->stream>>>(d, u); h->launches++; } });
  if (d.alpha_smooth) {
    if (h->cfg.nranks > 1 && !h->peers.empty()) {
      // indicator values of the elements behind the cut faces (one double per exchanged face), on the main stream
      if (!h->comm) return fail(TRIXIB200_ECOMM, "nranks > 1 but trixib200_comm_init was not called");
      LAUNCH(h, k_pack_alpha, d.nhalo_send, 128, 0, d);
      int rc = g_nccl.GroupStart(), rc1 = 0;
      size_t soff = 0, roff = 0;
      for (size_t k = 0; k < h->peers.size() && rc == 0; ++k) {
        const size_t sc = (size_t)h->peer_send_count[k], rcn = (size_t)h->peer_count[k];
        if (sc > 0 && !rc1) rc1 = g_nccl.Send(d.halo_alpha_send + soff, sc, NCCL_FLOAT64, h->peers[k], h->comm, h->stream);
        if (rcn > 0 && !rc1) rc1 = g_nccl.Recv(d.halo_alpha_recv + roff, rcn, NCCL_FLOAT64, h->peers[k], h->comm, h->stream);
        soff += sc; roff += rcn;
      }
      if (rc == 0) rc = g_nccl.GroupEnd();
      if (rc != 0 || rc1 != 0) return fail(TRIXIB200_ECOMM, "alpha halo exchange (ncclSend/ncclRecv) failed");
    }
    LAUNCH(h, k_alpha_smooth_interfaces, d.I, 256, 0, d);
    LAUNCH(h, k_alpha_smooth_mortars, d.M, 128, 0, d, h->nsmall);
  }
  return 0;
}
static int st_volume(trixib200_handle* h, double* du, const double* u) {
  Dev& d = h->d;
  if (d.volume_integral == TRIXIB200_VI_SHOCK_CAPTURING_HG)
    if (int rc = st_indicator(h, u)) return rc;
  TB_DISPATCH_EQ(h, LAUNCH(h, k_volume<Eq>, d.E * d.nn, 128, 0, d, du, u));
  return 0;
}
static int st_prolong_interfaces(trixib200_handle* h, const double* u) {
  Dev& d = h->d;
  TB_DISPATCH_ND(h, LAUNCH(h, k_prolong_interfaces<ND>, d.I * d.nf * d.nv, 256, 0, d, u));
  return 0;
}
static int st_interface_flux(trixib200_handle* h) {
  Dev& d = h->d;
  TB_DISPATCH_EQ(h, LAUNCH(h, k_interface_flux<Eq>, d.I * d.nf, 128, 0, d));
  return 0;
}
static int st_prolong_boundaries(trixib200_handle* h, const double* u) {
  Dev& d = h->d;
  TB_DISPATCH_ND(h, LAUNCH(h, k_prolong_boundaries<ND>, d.B * d.nf * d.nv, 256, 0, d, u));
  return 0;
}
static int st_boundary_flux(trixib200_handle* h, double t) {
  Dev& d = h->d;
  TB_DISPATCH_EQ(h, LAUNCH(h, k_boundary_flux<Eq>, d.B * d.nf, 128, 0, d, t));
  return 0;
}
static int st_prolong_mortars(trixib200_handle* h, const double* u) {
  Dev& d = h->d;
  if (d.ndim == 2) LAUNCH(h, k_prolong_mortars<2>, d.M * d.nf * d.nv, 128, 0, d, u);
  else if (d.ndim == 3) LAUNCH(h, k_prolong_mortars<3>, d.M * d.nf * d.nv, 128, 0, d, u);
  return 0;
}
static int st_mortar_flux(trixib200_handle* h) {
  Dev& d = h->d;
  if (d.ndim == 1 || d.M == 0) return 0;
  switch (h->cfg.equations) {
    case TRIXIB200_EQ_ADVECTION:
      if (d.ndim == 2) LAUNCH(h, k_mortar_flux<EqAdvection<2>>, d.M * d.nf * 2, 128, 0, d);
      else LAUNCH(h, k_mortar_flux<EqAdvection<3>>, d.M * d.nf * 4, 128, 0, d);
      break;
    case TRIXIB200_EQ_EULER:
      if (d.ndim == 2) LAUNCH(h, k_mortar_flux<EqEuler<2>>, d.M * d.nf * 2, 128, 0, d);
      else LAUNCH(h, k_mortar_flux<EqEuler<3>>, d.M * d.nf * 4, 128, 0, d);
      break;
    default: LAUNCH(h, k_mortar_flux<EqMhd3>, d.M * d.nf * 4, 128, 0, d); break;
  }
  if (d.ndim == 2) LAUNCH(h, k_mortar_to_elements<2>, d.M * d.nf * d.nv, 128, 0, d);
  else LAUNCH(h, k_mortar_to_elements<3>, d.M * d.nf * d.nv, 128, 0, d);
  return 0;
}
static int st_epilogue(trixib200_handle* h, double* du, const double* u, double t, int flags) {
  Dev& d = h->d;
  TB_DISPATCH_EQ(h, LAUNCH(h, k_epilogue<Eq>, d.E * d.nn, 128, 0, d, du, u, t, flags));
  return 0;
}

// halo exchange: pack + NCCL send/recv on the comm stream, beside the interior elements on the main stream
static int halo_begin(trixib200_handle* h, const double* u) {
  Dev& d = h->d;
  if (h->cfg.nranks == 1 || h->peers.empty()) return 0;
  if (!h->comm) return fail(TRIXIB200_ECOMM, "nranks > 1 but trixib200_comm_init was not called");
  // u is ready (and the previous rhs! is done with halo_recv) once the main stream reaches this point
  CUDA_TRY(cudaEventRecord(h->ev_pack, h->stream));
  CUDA_TRY(cudaStreamWaitEvent(h->comm_stream, h->ev_pack, 0));
  {
    // 128-thread blocks: they still fit next to the persistent fused kernel (registers), so packing does not delay
    // the interior launch (it used to sit in front of it on the main stream: 0.03-0.04 ms per rhs!)
    const int64_t n = d.nhalo_send * d.nf * d.nv;
    if (n > 0) {
      if (d.ndim == 1) k_pack_halo<1><<<nblk(n, 128), 128, 0, h->comm_stream>>>(d, u);
      else if (d.ndim == 2) k_pack_halo<2><<<nblk(n, 128), 128, 0, h->comm_stream>>>(d, u);
      else k_pack_halo<3><<<nblk(n, 128), 128, 0, h->comm_stream>>>(d, u);
      h->launches++;
    }
  }
  size_t per = (size_t)d.nv * d.nf;
  auto nccl_err = [](const char* what, int rc) {
    return fail(TRIXIB200_ECOMM, std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"));
  };
  int rc = g_nccl.GroupStart();
  if (rc != 0) return nccl_err("ncclGroupStart", rc);
  size_t soff = 0, roff = 0;
  int rc_first = 0;
  const char* what_first = nullptr;
  for (size_t k = 0; k < h->peers.size(); ++k) {
    const size_t scnt = per * (size_t)h->peer_send_count[k], rcnt = per * (size_t)h->peer_count[k];
    if (scnt > 0) {
      rc = g_nccl.Send(d.halo_send + soff, scnt, NCCL_FLOAT64, h->peers[k], h->comm, h->comm_stream);
      if (rc != 0 && !rc_first) { rc_first = rc; what_first = "ncclSend"; }
    }
    if (rcnt > 0) {
      rc = g_nccl.Recv((void*)(d.halo_recv + roff), rcnt, NCCL_FLOAT64, h->peers[k], h->comm, h->comm_stream);
      if (rc != 0 && !rc_first) { rc_first = rc; what_first = "ncclRecv"; }
    }
    soff += scnt; roff += rcnt;
  }
  rc = g_nccl.GroupEnd();       // always closed, also after a failed send / recv
  if (rc_first != 0) return nccl_err(what_first, rc_first);
  if (rc != 0) return nccl_err("ncclGroupEnd", rc);
  CUDA_TRY(cudaEventRecord(h->ev_halo, h->comm_stream));
  return 0;
}
static int halo_wait(trixib200_handle* h) {
  if (h->cfg.nranks == 1 || h->peers.empty()) return 0;
  CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_halo, 0));
  return 0;
}

static int rhs_staged(trixib200_handle* h, double* du, const double* u, double t) {
  if (int rc = halo_begin(h, u)) return rc;
  if (int rc = st_volume(h, du, u)) return rc;
  if (int rc = halo_wait(h)) return rc;
  st_prolong_interfaces(h, u);
  st_interface_flux(h);
  st_prolong_boundaries(h, u);
  st_boundary_flux(h, t);
  st_prolong_mortars(h, u);
  st_mortar_flux(h);
  st_epilogue(h, du, u, t, 7);
  return 0;
}

// can rhs! + 2N Runge-Kutta stage run as ONE launch (k_line6<..., RK = true>) on these vectors?
static bool rk_fusable(const trixib200_handle* h, const double* u_out, const double* u_in, const double* tmp) {
  return h->fused && h->line3d &&
         ((((uintptr_t)u_out) | ((uintptr_t)u_in) | ((uintptr_t)tmp)) & 15) == 0;
}

static int fused_launch_any(trixib200_handle* h, double* du, const double* u, double t, const int* elems, int64_t count,
                            const RkArgs* rk = nullptr, const P2PArgs* p2p = nullptr, const Dev* dev = nullptr) {
  if (count <= 0) return 0;
  const Dev& d = dev ? *dev : h->d;
  if (rk) {   // caller checked rk_fusable(): du is u_out here
    if (int rc = line6_launch_rk(h->cfg, d, h->line_ops, du, u, t, elems, count, h->stream, h->sm_count, *rk,
                                 p2p ? *p2p : P2PArgs{}))
      return fail(rc, "fused RK launch failed");
    h->launches++;
    return 0;
  }
  if (p2p) {  // caller checked that the line-owner kernel takes these vectors
    if (int rc = line6_launch(h->cfg, d, h->line_ops, du, u, t, elems, count, h->stream, h->sm_count, *p2p))
      return fail(rc, "fused launch failed");
    h->launches++;
    return 0;
  }
  const bool w3 = h->warp3d && ((((uintptr_t)u) | ((uintptr_t)du)) & 15) == 0;
  const bool l3 = w3 && h->line3d;
  int rc = l3 ? line6_launch(h->cfg, d, h->line_ops, du, u, t, elems, count, h->stream, h->sm_count)
         : w3 ? warp3d_launch(h->cfg, d, du, u, t, elems, count, h->stream, h->sm_count)
              : fused_launch(h->cfg, d, du, u, t, elems, count, h->stream, h->sm_count);
  if (rc) return fail(rc, "fused launch failed");
  h->launches++;
  return 0;
}

// TRIXIB200_TRACE=n: on the n-th multi-rank rhs! of a handle, time its pieces with CUDA events and print them
struct FusedTrace {
  cudaEvent_t e[8];
  bool made = false;
  int64_t calls = 0;
};
static FusedTrace g_trace;

static int rhs_fused(trixib200_handle* h, double* du, const double* u, double t, const RkArgs* rk = nullptr) {
  Dev& d = h->d;
  static const int trace_at = getenv("TRIXIB200_TRACE") ? atoi(getenv("TRIXIB200_TRACE")) : 0;
  const bool tracing = trace_at > 0 && h->cfg.nranks > 1 && ++g_trace.calls == trace_at;
  if (tracing) {
    if (!g_trace.made) { for (auto& e : g_trace.e) cudaEventCreate(&e); g_trace.made = true; }
    cudaDeviceSynchronize();
    cudaEventRecord(g_trace.e[0], h->stream);
  }
  // multi-GPU on the line-owner path with mapped peer buffers: ONE launch packs, exchanges and computes
  // (the decision must be the same on every rank: it only uses global facts)
  if (h->p2p.enabled && h->cfg.nranks > 1 && h->p2p.mesh_ok) {
    if ((((uintptr_t)u) | ((uintptr_t)du) | (rk ? (uintptr_t)rk->tmp : 0)) & 15)
      return fail(TRIXIB200_EINVAL, "multi-GPU rhs!: vectors must be 16-byte aligned");
    auto& P = h->p2p;
    const unsigned long long epoch = ++P.epoch;
    const int par = (int)(epoch & 1);
    Dev dl = d;
    dl.halo_recv = (const double*)(P.base + P2P_FLAG_BYTES) + (size_t)par * P.buf_doubles;
    P2PArgs a;
    a.npeers = (int)h->peers.size();
    a.first_cut_pair = (int)(h->n_interior / 2);
    a.epoch = epoch;
    a.peer_first = P.d_peer_first; a.peer_rank = P.d_peer_rank;
    a.peer_dst = P.d_peer_dst[par]; a.peer_flag = P.d_peer_flag;
    a.my_flags = (const unsigned long long*)P.base;
    a.done_counter = P.d_done;
    return fused_launch_any(h, du, u, t, P.d_elems_all, d.E, rk, &a, &dl);
  }
  if (int rc = halo_begin(h, u)) return rc;
  if (tracing) { cudaEventRecord(g_trace.e[1], h->stream); cudaEventRecord(g_trace.e[5], h->comm_stream); }
  if (d.volume_integral == TRIXIB200_VI_SHOCK_CAPTURING_HG)
    if (int rc = st_indicator(h, u)) return rc;
  // faces the fused kernel does not compute itself: boundary and mortar faces -> surface_flux_values
  if (d.B > 0) { st_prolong_boundaries(h, u); st_boundary_flux(h, t); }
  bool multi = h->cfg.nranks > 1 && !h->peers.empty();
  bool halo_here = false;
  if (d.M > 0) {
    // a mortar replicated on this rank may need face traces of elements on other ranks: they come with the halo
    if (multi) { if (int rc = halo_wait(h)) return rc; halo_here = true; }
    st_prolong_mortars(h, u); st_mortar_flux(h);
  }
  // (the line-owner / warp-per-element kernels need 16-byte aligned vectors; fused_launch_any checks)
  auto launch = [&](const int* elems, int64_t count) -> int { return fused_launch_any(h, du, u, t, elems, count, rk); };
  if (!multi || halo_here) {
    if (int rc = launch(nullptr, d.E)) return rc;
  } else {
    // (TRIXIB200_TRACE on 2 GPUs, level 7: pack 0.03-0.04 ms, exchange complete 0.12 ms after the start -- it runs
    // beside the interior launch, 2.33 ms -- no wait for the halo, cut elements 0.10 ms. Leaving SMs free for the
    // NCCL kernels only costs their share of the interior throughput.)
    if (int rc = launch(h->d_elems_interior, h->n_interior)) return rc;
    if (tracing) cudaEventRecord(g_trace.e[2], h->stream);
    if (int rc = halo_wait(h)) return rc;
    if (tracing) cudaEventRecord(g_trace.e[3], h->stream);
    if (int rc = launch(h->d_elems_halo, h->n_halo_elems)) return rc;
    if (tracing) {
      cudaEventRecord(g_trace.e[4], h->stream);
      cudaDeviceSynchronize();
      float pack, interior, wait, cut, exch_end;
      cudaEventElapsedTime(&pack, g_trace.e[0], g_trace.e[1]);
      cudaEventElapsedTime(&interior, g_trace.e[1], g_trace.e[2]);
      cudaEventElapsedTime(&wait, g_trace.e[2], g_trace.e[3]);
      cudaEventElapsedTime(&cut, g_trace.e[3], g_trace.e[4]);
      cudaEventElapsedTime(&exch_end, g_trace.e[0], g_trace.e[5]);
      fprintf(stderr, "[trixib200 trace rank %d] pack %.3f ms | interior %.3f | wait for halo %.3f | cut elements %.3f | "
              "exchange done %.3f ms after start | interior %lld cut %lld elements\n", h->cfg.rank, pack, interior, wait, cut,
              exch_end, (long long)h->n_interior, (long long)h->n_halo_elems);
    }
  }
  return 0;
}

extern "C" int trixib200_rhs(trixib200_handle* h, double* du, const double* u, double t) {
  if (!h || !du || !u) return fail(TRIXIB200_EINVAL, "null argument");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  int rc = h->fused ? rhs_fused(h, du, u, t) : rhs_staged(h, du, u, t);
  if (rc) return rc;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------- rhs on host vectors
// Slab pipeline of trixib200_rhs_host. The local elements are grouped into slabs by their last-dimension coordinate
// (TRIXIB200_HOST_SLABS of them, default 16: thin layers of the mesh). Slabs are uploaded in order -- one copy per
// contiguous run of the Morton order inside the slab --, a slab is computed with ONE launch over its element list as
// soon as the slabs holding all its face neighbours have landed, and its du goes back right after: upload, kernels and
// download overlap (PCIe is full duplex) instead of running back to back. Thinner slabs shorten the fill and drain of
// the pipeline (tools/host_pipe_model.py); round 1 launched one kernel per 4096-element chunk, which made thin slabs
// cost more in launches than they saved. Used when one fused launch is the whole rhs! (single rank, no boundary /
// mortar faces, no shock-capturing indicator pass); otherwise the plain sequence.
static int build_host_pipe(trixib200_handle* h) {
  auto& P = h->pipe;
  P.built = true;
  const Dev& d = h->d;
  P.usable = h->fused && h->cfg.nranks == 1 && d.B == 0 && d.M == 0 &&
             d.volume_integral != TRIXIB200_VI_SHOCK_CAPTURING_HG && !h->face_nbr_host.empty();
  const int64_t E = d.E;
  int64_t ce = 4096;   // the pipeline pays from ~4 such chunks on; TRIXIB200_HOST_CHUNK lowers the threshold (tests)
  const char* env = getenv("TRIXIB200_HOST_CHUNK");
  if (env && atoll(env) >= 64) ce = atoll(env);
  P.usable = P.usable && E >= 4 * ce;
  if (!P.usable) return 0;
  P.ce = ce;
  int want = 16;
  const char* es = getenv("TRIXIB200_HOST_SLABS");
  if (es && atoi(es) >= 2) want = atoi(es);
  // distinct layers of the last coordinate -> slab of every element
  std::vector<double> keys(h->chunk_key_src);
  std::sort(keys.begin(), keys.end());
  keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
  const int64_t nlayers = (int64_t)keys.size();
  const int64_t S = std::max<int64_t>(2, std::min<int64_t>(want, nlayers));
  P.nchunks = S;
  std::vector<int> slab((size_t)E);
  for (int64_t e = 0; e < E; ++e) {
    const int64_t layer = std::lower_bound(keys.begin(), keys.end(), h->chunk_key_src[e]) - keys.begin();
    slab[e] = (int)(layer * S / nlayers);
  }
  // element list per slab (ascending element id), contiguous runs inside it
  std::vector<int64_t> cnt(S + 1, 0);
  for (int64_t e = 0; e < E; ++e) cnt[slab[e] + 1]++;
  for (int64_t q = 0; q < S; ++q) cnt[q + 1] += cnt[q];
  P.slab_off.assign(cnt.begin(), cnt.end());
  std::vector<int> elems((size_t)E);
  {
    std::vector<int64_t> pos(cnt.begin(), cnt.end() - 1);
    for (int64_t e = 0; e < E; ++e) elems[pos[slab[e]]++] = (int)e;
  }
  P.runs.assign(S, {});
  for (int64_t q = 0; q < S; ++q)
    for (int64_t k = P.slab_off[q]; k < P.slab_off[q + 1];) {
      int64_t k1 = k + 1;
      while (k1 < P.slab_off[q + 1] && elems[k1] == elems[k1 - 1] + 1) ++k1;
      P.runs[q].push_back({(int64_t)elems[k], k1 - k});
      k = k1;
    }
  // a slab can be computed once every slab that holds a face neighbour of one of its elements has landed
  const int nf = 2 * d.ndim;
  P.ready.assign(S, {});
  for (int64_t q = 0; q < S; ++q) {
    int r = (int)q;
    for (int64_t k = P.slab_off[q]; k < P.slab_off[q + 1]; ++k)
      for (int f = 0; f < nf; ++f) {
        const int code = h->face_nbr_host[(size_t)elems[k] * nf + f];
        if (code >= 0) r = std::max(r, slab[code]);
      }
    P.ready[r].push_back((int)q);
  }
  CUDA_TRY(cudaStreamCreateWithFlags(&P.s_in, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&P.s_out, cudaStreamNonBlocking));
  P.ev_in.resize(S); P.ev_done.resize(S);
  for (int64_t c = 0; c < S; ++c) {
    CUDA_TRY(cudaEventCreateWithFlags(&P.ev_in[c], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&P.ev_done[c], cudaEventDisableTiming));
  }
  if (int rc = upload(h, elems, &P.d_iota)) return rc;
  return 0;
}

extern "C" int trixib200_rhs_host(trixib200_handle* h, double* du_host, const double* u_host, double t) {
  if (!h || !du_host || !u_host) return fail(TRIXIB200_EINVAL, "null argument");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  const Dev& d = h->d;
  const size_t n = (size_t)d.E * d.nn * d.nv;
  if (!h->host_u) {
    if (int rc = dalloc(h, n, &h->host_u, false)) return rc;
    if (int rc = dalloc(h, n, &h->host_du, false)) return rc;
  }
  if (!h->pipe.built)
    if (int rc = build_host_pipe(h)) return rc;
  auto& P = h->pipe;
  if (!P.usable) {
    CUDA_TRY(cudaMemcpyAsync(h->host_u, u_host, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if (int rc = trixib200_rhs(h, h->host_du, h->host_u, t)) return rc;
    CUDA_TRY(cudaMemcpyAsync(du_host, h->host_du, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return 0;
  }
  // the previous work on the handle's stream may still read host_u / write host_du
  CUDA_TRY(cudaEventRecord(h->ev_pack, h->stream));
  CUDA_TRY(cudaStreamWaitEvent(P.s_in, h->ev_pack, 0));
  const size_t per = (size_t)d.nn * d.nv;
  // Optional (TRIXIB200_HOST_DIRECT=1) for page-locked du: the kernels store du straight into host memory (posted
  // PCIe writes, 16-byte coalesced) instead of a copy stage. Measured on B200 / PCIe Gen5 at level 7: 144 ms per
  // rhs! against 140 ms with the DMA download -- both sit on the duplex PCIe limit (2 x 5.37 GB), so it is off.
  double* du_direct = nullptr;
  {
    static const bool allow = getenv("TRIXIB200_HOST_DIRECT") && atoi(getenv("TRIXIB200_HOST_DIRECT")) == 1;
    cudaPointerAttributes at;
    if (allow && cudaPointerGetAttributes(&at, du_host) == cudaSuccess && at.type == cudaMemoryTypeHost &&
        at.devicePointer && (((uintptr_t)at.devicePointer) & 15) == 0)
      du_direct = (double*)at.devicePointer;
    cudaGetLastError();
  }
  for (int64_t i = 0; i < P.nchunks; ++i) {
    for (const auto& run : P.runs[i])
      CUDA_TRY(cudaMemcpyAsync(h->host_u + per * run.first, u_host + per * run.first, per * run.count * sizeof(double),
                               cudaMemcpyHostToDevice, P.s_in));
    CUDA_TRY(cudaEventRecord(P.ev_in[i], P.s_in));
    if (P.ready[i].empty()) continue;
    CUDA_TRY(cudaStreamWaitEvent(h->stream, P.ev_in[i], 0));
    for (int c : P.ready[i]) {
      if (int rc = fused_launch_any(h, du_direct ? du_direct : h->host_du, h->host_u, t, P.d_iota + P.slab_off[c],
                                    P.slab_off[c + 1] - P.slab_off[c])) return rc;
      if (du_direct) continue;
      CUDA_TRY(cudaEventRecord(P.ev_done[c], h->stream));
      CUDA_TRY(cudaStreamWaitEvent(P.s_out, P.ev_done[c], 0));
      for (const auto& run : P.runs[c])
        CUDA_TRY(cudaMemcpyAsync(du_host + per * run.first, h->host_du + per * run.first,
                                 per * run.count * sizeof(double), cudaMemcpyDeviceToHost, P.s_out));
    }
  }
  CUDA_TRY(cudaStreamSynchronize(P.s_out));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  CUDA_TRY(cudaGetLastError());
  return 0;
}
extern "C" int trixib200_host_register(trixib200_handle* h, double* host, int64_t n) {
  if (!h || !host || n < 0) return fail(TRIXIB200_EINVAL, "bad argument");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  CUDA_TRY(cudaHostRegister(host, (size_t)n * sizeof(double), cudaHostRegisterDefault));
  return 0;
}
extern "C" int trixib200_host_unregister(trixib200_handle* h, double* host) {
  if (!h || !host) return fail(TRIXIB200_EINVAL, "bad argument");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  CUDA_TRY(cudaHostUnregister(host));
  return 0;
}

extern "C" int trixib200_stage(trixib200_handle* h, const char* stage, double* du, const double* u, double t) {
  if (!h || !stage) return fail(TRIXIB200_EINVAL, "null argument");
  if (h->fused) return fail(TRIXIB200_EINVAL, "per-stage entry points need a handle created with TRIXIB200_FLAG_STAGED_ONLY");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  std::string n(stage);
  Dev& d = h->d;
  if (n == "reset_du") CUDA_TRY(cudaMemsetAsync(du, 0, sizeof(double) * d.E * d.nn * d.nv, h->stream));
  else if (n == "calc_volume_integral") { if (int rc = st_volume(h, du, u)) return rc; }
  else if (n == "prolong2interfaces") {
    if (int rc = halo_begin(h, u)) return rc;
    if (int rc = halo_wait(h)) return rc;
    st_prolong_interfaces(h, u);
  }
  else if (n == "calc_interface_flux") st_interface_flux(h);
  else if (n == "prolong2boundaries") st_prolong_boundaries(h, u);
  else if (n == "calc_boundary_flux") st_boundary_flux(h, t);
  else if (n == "prolong2mortars") st_prolong_mortars(h, u);
  else if (n == "calc_mortar_flux") st_mortar_flux(h);
  else if (n == "calc_surface_integral") st_epilogue(h, du, u, t, 1);
  else if (n == "apply_jacobian") st_epilogue(h, du, u, t, 2);
  else if (n == "calc_sources") st_epilogue(h, du, u, t, 4);
  else if (n == "calc_indicator") { if (int rc = st_indicator(h, u)) return rc; }
  else return fail(TRIXIB200_EINVAL, "unknown stage " + n);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return 0;
}

static bool cache_array(const trixib200_handle* h, const std::string& n, const double** p, int64_t* len) {
  const Dev& d = h->d;
  int64_t face = (int64_t)d.nv * d.nf;
  if (n == "interfaces.u") { *p = d.interfaces_u; *len = 2 * face * d.I; return d.interfaces_u != nullptr; }
  if (n == "boundaries.u") { *p = d.boundaries_u; *len = 2 * face * d.B; return true; }
  if (n == "surface_flux_values") { *p = d.sfv; *len = face * 2 * d.ndim * d.E; return d.sfv != nullptr; }
  if (n == "alpha") { *p = d.alpha; *len = d.E; return true; }
  const char* m3[4] = {"mortars.u_upper_left", "mortars.u_upper_right", "mortars.u_lower_left", "mortars.u_lower_right"};
  const char* m2[2] = {"mortars.u_upper", "mortars.u_lower"};
  for (int q = 0; q < 4; ++q) if (d.ndim == 3 && n == m3[q]) { *p = d.mortar_u[q]; *len = 2 * face * d.M; return true; }
  for (int q = 0; q < 2; ++q) if (d.ndim == 2 && n == m2[q]) { *p = d.mortar_u[q]; *len = 2 * face * d.M; return true; }
  return false;
}
extern "C" int64_t trixib200_cache_len(const trixib200_handle* h, const char* name) {
  const double* p; int64_t len;
  if (!h || !name || !cache_array(h, name, &p, &len)) return -1;
  return len;
}
extern "C" int trixib200_cache_get(trixib200_handle* h, const char* name, double* out, int64_t n) {
  const double* p; int64_t len;
  if (!h || !name || !cache_array(h, name, &p, &len)) return fail(TRIXIB200_EINVAL, "unknown or unavailable cache array");
  if (len != n) return fail(TRIXIB200_EINVAL, "length mismatch");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (n > 0) CUDA_TRY(cudaMemcpy(out, p, sizeof(double) * n, cudaMemcpyDeviceToHost));
  return 0;
}

// ---------------------------------------------------------------------------------------------- max_dt
extern "C" int trixib200_max_dt(trixib200_handle* h, const double* u, double t, double* out_host) {
  (void)t;
  if (!h || !u || !out_host) return fail(TRIXIB200_EINVAL, "null argument");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  Dev& d = h->d;
  // nextfloat(0.0): avoids a division by zero if the speed vanishes (reference stepsize_dg_3d.jl:22-24)
  double init = 4.9406564584124654e-324;
  CUDA_TRY(cudaMemcpyAsync(h->d_scalar, &init, sizeof(double), cudaMemcpyHostToDevice, h->stream));
  int64_t warps = std::min<int64_t>(d.E, (int64_t)h->sm_count * 64);
  TB_DISPATCH_EQ(h, LAUNCH(h, k_max_dt<Eq>, warps * 32, 256, 0, d, u, h->d_scalar));
  if (h->cfg.nranks > 1) {
    if (!h->comm) return fail(TRIXIB200_ECOMM, "nranks > 1 but trixib200_comm_init was not called");
    int rc = g_nccl.AllReduce(h->d_scalar, h->d_scalar, 1, NCCL_FLOAT64, NCCL_MAX, h->comm, h->stream);
    if (rc != 0) return fail(TRIXIB200_ECOMM, "ncclAllReduce failed");
  }
  double m = 0;
  CUDA_TRY(cudaMemcpyAsync(&m, h->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  *out_host = 2 / (d.N * m);
  return 0;
}

// ---------------------------------------------------------------------------------------------- memory etc.
extern "C" int trixib200_alloc(trixib200_handle* h, int64_t n, double** out) {
  if (!h || !out || n < 0) return fail(TRIXIB200_EINVAL, "bad argument");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  void* p = nullptr;
  if (cudaMalloc(&p, std::max<int64_t>(n, 1) * sizeof(double)) != cudaSuccess) return fail(TRIXIB200_ENOMEM, "cudaMalloc failed");
  *out = (double*)p;
  return 0;
}
extern "C" int trixib200_free(trixib200_handle* h, double* p) {
  if (!h) return fail(TRIXIB200_EINVAL, "null handle");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  CUDA_TRY(cudaFree(p));
  return 0;
}
extern "C" int trixib200_upload(trixib200_handle* h, double* dst, const double* src, int64_t n) {
  if (!h) return fail(TRIXIB200_EINVAL, "null handle");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  CUDA_TRY(cudaMemcpyAsync(dst, src, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return 0;
}
extern "C" int trixib200_download(trixib200_handle* h, double* dst, const double* src, int64_t n) {
  if (!h) return fail(TRIXIB200_EINVAL, "null handle");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  CUDA_TRY(cudaMemcpyAsync(dst, src, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return 0;
}
extern "C" int trixib200_sync(trixib200_handle* h) {
  if (!h) return fail(TRIXIB200_EINVAL, "null handle");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->comm_stream));
  return 0;
}
extern "C" int trixib200_set_stream(trixib200_handle* h, int64_t stream) {
  if (!h) return fail(TRIXIB200_EINVAL, "null handle");
  if ((cudaStream_t)(intptr_t)stream == h->stream) return 0;
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (h->owns_stream) { cudaStreamDestroy(h->stream); h->owns_stream = false; }
  h->stream = (cudaStream_t)(intptr_t)stream;
  return 0;
}
extern "C" int64_t trixib200_stream(const trixib200_handle* h) { return h ? (int64_t)(intptr_t)h->stream : 0; }
extern "C" int64_t trixib200_launch_count(const trixib200_handle* h) { return h ? h->launches : 0; }

extern "C" int trixib200_fill_initial_condition(trixib200_handle* h, double* u, double t) {
  if (!h || !u) return fail(TRIXIB200_EINVAL, "null argument");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  Dev& d = h->d;
  if (!ic_supported(h->cfg)) return fail(TRIXIB200_EUNSUPPORTED, "fill_initial_condition: the initial condition is not enumerated");
  if (!d.node_coords && !d.centers) return fail(TRIXIB200_EINVAL, "fill_initial_condition needs node_coordinates or cell_centers");
  TB_DISPATCH_EQ(h, LAUNCH(h, k_fill_ic<Eq>, d.E * d.nn, 128, 0, d, u, t));
  CUDA_TRY(cudaGetLastError());
  return 0;
}

extern "C" int trixib200_rk2n_update(trixib200_handle* h, double* u, double* tmp, const double* du, double a,
                                     double b, double dt) {
  if (!h || !u || !tmp || !du) return fail(TRIXIB200_EINVAL, "null argument");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  int64_t n = h->d.E * h->d.nn * h->d.nv;
  if (n > 0) {
    int64_t blocks = std::min<int64_t>((n + 255) / 256, (int64_t)h->sm_count * 16);
    k_rk2n_update<<<(unsigned)blocks, 256, 0, h->stream>>>(u, tmp, du, a, b, dt, n);
    h->launches++;
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// out-of-place 2N stage for the kernel families without the fused epilogue
__global__ void k_rk2n_stage_oop(double* __restrict__ u_out, const double* __restrict__ u_in, double* __restrict__ tmp,
                                 const double* __restrict__ du, double a, double b, double dt, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const double tn = a != 0.0 ? fma(a, tmp[i], dt * du[i]) : dt * du[i];
    tmp[i] = tn;
    u_out[i] = fma(b, tn, u_in[i]);
  }
}

extern "C" int trixib200_rk2n_stage(trixib200_handle* h, double* u_out, const double* u_in, double* tmp, double t,
                                    double a, double b, double dt) {
  if (!h || !u_out || !u_in || !tmp) return fail(TRIXIB200_EINVAL, "null argument");
  if (u_out == u_in) return fail(TRIXIB200_EINVAL, "rk2n_stage: u_out must not alias u_in (neighbours read u_in)");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  if (rk_fusable(h, u_out, u_in, tmp)) {
    const RkArgs rk{tmp, a, b, dt};
    if (int rc = rhs_fused(h, u_out, u_in, t, &rk)) return rc;
    CUDA_TRY(cudaGetLastError());
    return 0;
  }
  const int64_t n = h->d.E * h->d.nn * h->d.nv;
  if (!h->du_scratch)
    if (int rc = dalloc(h, (size_t)n, &h->du_scratch, false)) return rc;
  if (int rc = trixib200_rhs(h, h->du_scratch, u_in, t)) return rc;
  if (n > 0) {
    const int64_t blocks = std::min<int64_t>((n + 255) / 256, (int64_t)h->sm_count * 16);
    k_rk2n_stage_oop<<<(unsigned)blocks, 256, 0, h->stream>>>(u_out, u_in, tmp, h->du_scratch, a, b, dt, n);
    h->launches++;
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// CarpenterKennedy2N54 coefficients (Trixi / OrdinaryDiffEq `CarpenterKennedy2N54`)
static const double CK_A[5] = {0.0, -567301805773.0 / 1357537059087.0, -2404267990393.0 / 2016746695238.0,
                               -3550918686646.0 / 2091501179385.0, -1275806237668.0 / 842570457699.0};
static const double CK_B[5] = {1432997174477.0 / 9575080441755.0, 5161836677717.0 / 13612068292357.0,
                               1720146321549.0 / 2090206949498.0, 3134564353537.0 / 4481467310338.0,
                               2277821191437.0 / 14882151754819.0};
static const double CK_C[5] = {0.0, 1432997174477.0 / 9575080441755.0, 2526269341429.0 / 6820363962896.0,
                               2006345519317.0 / 3224310063776.0, 2802321613138.0 / 2924317926251.0};

extern "C" int trixib200_rk2n_step_ck54(trixib200_handle* h, double* u, double* u_alt, double* tmp, double t, double dt,
                                        int* result_in_alt) {
  if (!h || !u || !u_alt || !tmp) return fail(TRIXIB200_EINVAL, "null argument");
  double* cur = u;
  double* nxt = u_alt;
  for (int s = 0; s < 5; ++s) {
    if (int rc = trixib200_rk2n_stage(h, nxt, cur, tmp, t + CK_C[s] * dt, CK_A[s], CK_B[s], dt)) return rc;
    std::swap(cur, nxt);
  }
  if (result_in_alt) *result_in_alt = (cur == u_alt) ? 1 : 0;
  return 0;
}

// ---------------------------------------------------------------------------------------------- analysis
constexpr int AN_GRID = 1024;   // CTAs (= partials per variable) of the analysis kernels
static int analysis_buffers(trixib200_handle* h) {
  if (h->an_buf) return 0;
  return dalloc(h, (size_t)AN_GRID * 2 * AN_MAXV + 16 * MAXN + 64, &h->an_buf, false);
}
// sum / max of the local values over the ranks (in place, host vectors of n <= AN_MAXV doubles)
static int analysis_allreduce(trixib200_handle* h, double* v, int n, bool max) {
  if (h->cfg.nranks == 1) return 0;
  if (!h->comm) return fail(TRIXIB200_ECOMM, "nranks > 1 but trixib200_comm_init was not called");
  CUDA_TRY(cudaMemcpyAsync(h->an_buf, v, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  if (g_nccl.AllReduce(h->an_buf, h->an_buf, n, NCCL_FLOAT64, max ? NCCL_MAX : NCCL_SUM, h->comm, h->stream) != 0)
    return fail(TRIXIB200_ECOMM, "ncclAllReduce failed");
  CUDA_TRY(cudaMemcpyAsync(v, h->an_buf, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return 0;
}

extern "C" int trixib200_calc_error_norms(trixib200_handle* h, const double* u, double t, int32_t n_analysis,
                                          const double* vandermonde, const double* weights, double total_volume,
                                          double* l2_out, double* linf_out) {
  if (!h || !u || !vandermonde || !weights || !l2_out || !linf_out) return fail(TRIXIB200_EINVAL, "null argument");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  Dev& d = h->d;
  const int NA = n_analysis, nv = d.nv;
  if (NA < d.N || NA > 2 * MAXN) return fail(TRIXIB200_EINVAL, "calc_error_norms: nnodes <= n_analysis <= 16");
  if (!ic_supported(h->cfg)) return fail(TRIXIB200_EUNSUPPORTED, "calc_error_norms: the initial condition is not enumerated");
  if (!d.node_coords && !d.centers) return fail(TRIXIB200_EINVAL, "calc_error_norms needs node_coordinates or cell_centers");
  if (!(total_volume > 0)) return fail(TRIXIB200_EINVAL, "calc_error_norms: total_volume must be positive");
  const size_t smem = analysis_smem_doubles(d.ndim, d.N, NA, nv) * sizeof(double);
  if (smem > 200 * 1024) return fail(TRIXIB200_EUNSUPPORTED, "calc_error_norms: analysis tile exceeds shared memory");
  if (int rc = analysis_buffers(h)) return rc;
  double* dV = h->an_buf + (size_t)AN_GRID * 2 * AN_MAXV;
  double* dw = dV + 16 * MAXN;
  CUDA_TRY(cudaMemcpyAsync(dV, vandermonde, (size_t)NA * d.N * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaMemcpyAsync(dw, weights, (size_t)NA * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  const int grid = (int)std::min<int64_t>(std::max<int64_t>(d.E, 1), AN_GRID);
  TB_DISPATCH_EQ(h, {
    auto kern = k_error_norms<Eq>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return fail(TRIXIB200_ECUDA, "calc_error_norms: shared memory opt-in failed");
    kern<<<grid, AN_THREADS, smem, h->stream>>>(d, u, t, NA, dV, dw, h->an_buf);
    h->launches++;
  });
  CUDA_TRY(cudaGetLastError());
  std::vector<double> part((size_t)grid * 2 * nv);
  CUDA_TRY(cudaMemcpyAsync(part.data(), h->an_buf, part.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  double l2[AN_MAXV], li[AN_MAXV];
  for (int v = 0; v < nv; ++v) { l2[v] = 0; li[v] = 0; }
  for (int b = 0; b < grid; ++b)
    for (int v = 0; v < nv; ++v) {
      l2[v] += part[(2 * (size_t)b + 0) * nv + v];
      li[v] = std::max(li[v], part[(2 * (size_t)b + 1) * nv + v]);
    }
  if (int rc = analysis_allreduce(h, l2, nv, false)) return rc;
  if (int rc = analysis_allreduce(h, li, nv, true)) return rc;
  for (int v = 0; v < nv; ++v) { l2_out[v] = std::sqrt(l2[v] / total_volume); linf_out[v] = li[v]; }
  return 0;
}

extern "C" int trixib200_integrate(trixib200_handle* h, const double* u, int32_t normalize, double total_volume,
                                   double* out) {
  if (!h || !u || !out) return fail(TRIXIB200_EINVAL, "null argument");
  if (normalize && !(total_volume > 0)) return fail(TRIXIB200_EINVAL, "integrate: total_volume must be positive");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  Dev& d = h->d;
  const int nv = d.nv;
  if (int rc = analysis_buffers(h)) return rc;
  const int grid = (int)std::min<int64_t>(std::max<int64_t>(d.E, 1), AN_GRID);
  TB_DISPATCH_EQ(h, { k_integrate<Eq><<<grid, AN_THREADS, 0, h->stream>>>(d, u, h->an_buf); h->launches++; });
  CUDA_TRY(cudaGetLastError());
  std::vector<double> part((size_t)grid * nv);
  CUDA_TRY(cudaMemcpyAsync(part.data(), h->an_buf, part.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  double acc[AN_MAXV];
  for (int v = 0; v < nv; ++v) acc[v] = 0;
  for (int b = 0; b < grid; ++b)
    for (int v = 0; v < nv; ++v) acc[v] += part[(size_t)b * nv + v];
  if (int rc = analysis_allreduce(h, acc, nv, false)) return rc;
  for (int v = 0; v < nv; ++v) out[v] = normalize ? acc[v] / total_volume : acc[v];
  return 0;
}

extern "C" int trixib200_time_rhs(trixib200_handle* h, double* du, const double* u, double t, int reps, float* ms) {
  if (!h || !ms || reps < 1) return fail(TRIXIB200_EINVAL, "bad argument");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->comm_stream));
  CUDA_TRY(cudaEventRecord(h->ev_a, h->stream));
  for (int i = 0; i < reps; ++i)
    if (int rc = trixib200_rhs(h, du, u, t)) return rc;
  CUDA_TRY(cudaEventRecord(h->ev_b, h->stream));
  CUDA_TRY(cudaEventSynchronize(h->ev_b));
  CUDA_TRY(cudaEventElapsedTime(ms, h->ev_a, h->ev_b));
  return 0;
}

// ---------------------------------------------------------------------------------------------- comm
extern "C" int trixib200_comm_unique_id(char* id128) {
  if (!id128) return fail(TRIXIB200_EINVAL, "null argument");
  if (!g_nccl.load()) return fail(TRIXIB200_ECOMM, "libnccl.so.2 not found");
  ncclUniqueId id;
  int rc = g_nccl.GetUniqueId(&id);
  if (rc != 0) return fail(TRIXIB200_ECOMM, "ncclGetUniqueId failed");
  std::memcpy(id128, id.internal, 128);
  return 0;
}
// In-kernel halo exchange: map every peer's receive buffer + flag words with CUDA IPC (one process per GPU, one box).
// All ranks take the same decision (two collective agreements); on any failure the NCCL send/recv path stays.
struct P2PRecord {
  cudaIpcMemHandle_t handle;
  int64_t ok, nhalo;
  int64_t slot_off[32];                   // first halo slot of rank r in THIS rank's receive buffer (-1: not a peer)
};
static int p2p_agree(trixib200_handle* h, bool mine, bool* all) {
  double v = mine ? 0.0 : 1.0, out = 0;   // max over ranks of "failed"
  CUDA_TRY(cudaMemcpyAsync(h->d_scalar, &v, sizeof(double), cudaMemcpyHostToDevice, h->stream));
  if (g_nccl.AllReduce(h->d_scalar, h->d_scalar, 1, NCCL_FLOAT64, NCCL_MAX, h->comm, h->stream) != 0)
    return fail(TRIXIB200_ECOMM, "ncclAllReduce failed (p2p agreement)");
  CUDA_TRY(cudaMemcpyAsync(&out, h->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  *all = out == 0.0;
  return 0;
}
static int p2p_setup(trixib200_handle* h) {
  auto& P = h->p2p;
  Dev& d = h->d;
  const int nr = h->cfg.nranks, me = h->cfg.rank;
  const size_t per = (size_t)d.nv * d.nf;
  const char* env = getenv("TRIXIB200_HALO");
  bool ok = true;
  if (env && std::string(env) == "nccl") { ok = false; P.why = "TRIXIB200_HALO=nccl"; }
  else if (!(h->fused && h->line3d)) { ok = false; P.why = "only the line-owner kernel path packs in-kernel"; }
  else if (!P.mesh_ok) { ok = false; P.why = "meshes with boundary or mortar faces exchange through NCCL send/recv"; }
  else if (nr > 32) { ok = false; P.why = "more than 32 ranks"; }
  else if (!g_nccl.AllGather) { ok = false; P.why = "ncclAllGather not found"; }
  P2PRecord mine;
  std::memset(&mine, 0, sizeof(mine));
  if (ok) {
    P.buf_doubles = std::max<size_t>(per * (size_t)d.nhalo_recv, 1);
    const size_t bytes = P2P_FLAG_BYTES + 2 * P.buf_doubles * sizeof(double);
    if (cudaMalloc((void**)&P.base, bytes) != cudaSuccess || cudaMemset(P.base, 0, bytes) != cudaSuccess ||
        cudaIpcGetMemHandle(&mine.handle, P.base) != cudaSuccess) {
      cudaGetLastError();
      ok = false; P.why = "cudaMalloc / cudaIpcGetMemHandle failed";
    }
  }
  mine.ok = ok ? 1 : 0;
  mine.nhalo = d.nhalo_recv;
  for (int r = 0; r < 32; ++r) mine.slot_off[r] = r < nr ? P.slot_off[r] : -1;
  // every rank's record to every rank
  std::vector<P2PRecord> recs(nr);
  {
    unsigned char* dbuf = nullptr;
    CUDA_TRY(cudaMalloc((void**)&dbuf, sizeof(P2PRecord) * nr));
    CUDA_TRY(cudaMemcpyAsync(dbuf + sizeof(P2PRecord) * me, &mine, sizeof(P2PRecord), cudaMemcpyHostToDevice, h->stream));
    int rc = g_nccl.AllGather ? g_nccl.AllGather(dbuf + sizeof(P2PRecord) * me, dbuf, sizeof(P2PRecord), /*ncclChar*/ 0,
                                                 h->comm, h->stream) : -1;
    if (rc == 0) {
      CUDA_TRY(cudaMemcpyAsync(recs.data(), dbuf, sizeof(P2PRecord) * nr, cudaMemcpyDeviceToHost, h->stream));
      CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
    cudaFree(dbuf);
    if (rc != 0) { P.why = "ncclAllGather failed"; return 0; }    // collective failed everywhere alike: stay on NCCL p2p
  }
  for (int r = 0; r < nr; ++r)
    if (!recs[r].ok) { if (ok) P.why = "a peer rank could not set up the in-kernel halo exchange"; ok = false; }
  // open the peers' allocations
  const size_t np = h->peers.size();
  std::vector<double*> dst[2];
  std::vector<unsigned long long*> flag;
  if (ok) {
    P.peer_base.assign(np, nullptr);
    for (size_t k = 0; k < np && ok; ++k) {
      const int pr = h->peers[k];
      if (recs[pr].slot_off[me] < 0) { ok = false; P.why = "peer tables disagree"; break; }
      void* pb = nullptr;
      if (cudaIpcOpenMemHandle(&pb, recs[pr].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        ok = false; P.why = "cudaIpcOpenMemHandle failed (no peer access between these GPUs?)";
        break;
      }
      P.peer_base[k] = pb;
      const size_t rbuf = std::max<size_t>(per * (size_t)recs[pr].nhalo, 1);
      double* b0 = (double*)((unsigned char*)pb + P2P_FLAG_BYTES);
      dst[0].push_back(b0 + per * (size_t)recs[pr].slot_off[me]);
      dst[1].push_back(b0 + rbuf + per * (size_t)recs[pr].slot_off[me]);
      flag.push_back((unsigned long long*)pb + me);
    }
  }
  bool all = false;
  if (int rc = p2p_agree(h, ok, &all)) return rc;
  if (!all) {
    if (ok) P.why = "a peer rank could not map the buffers";
    for (void*& pb : P.peer_base) if (pb) { cudaIpcCloseMemHandle(pb); pb = nullptr; }
    P.peer_base.clear();
    return 0;
  }
  std::vector<int> first(np + 1, 0), ranks(np, 0);
  for (size_t k = 0; k < np; ++k) { first[k + 1] = first[k] + (int)h->peer_send_count[k]; ranks[k] = h->peers[k]; }
  if (int rc = upload(h, first, &P.d_peer_first)) return rc;
  if (int rc = upload(h, ranks, &P.d_peer_rank)) return rc;
  if (int rc = upload(h, dst[0], &P.d_peer_dst[0])) return rc;
  if (int rc = upload(h, dst[1], &P.d_peer_dst[1])) return rc;
  if (int rc = upload(h, flag, &P.d_peer_flag)) return rc;
  std::vector<unsigned int> zero(1, 0u);
  if (int rc = upload(h, zero, &P.d_done)) return rc;
  P.enabled = true;
  P.why = "";
  return 0;
}

extern "C" int trixib200_comm_init(trixib200_handle* h, const char* id128) {
  if (!h || !id128) return fail(TRIXIB200_EINVAL, "null argument");
  if (h->cfg.nranks == 1) return 0;
  if (!g_nccl.load()) return fail(TRIXIB200_ECOMM, "libnccl.so.2 not found");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  ncclUniqueId id;
  std::memcpy(id.internal, id128, 128);
  int rc = g_nccl.CommInitRank(&h->comm, h->cfg.nranks, id, h->cfg.rank);
  if (rc != 0) return fail(TRIXIB200_ECOMM, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"));
  return p2p_setup(h);
}
