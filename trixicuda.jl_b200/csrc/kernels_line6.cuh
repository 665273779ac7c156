// Line-owner fused rhs! kernel, generation 6: the same ownership as k_line3d (a half-warp owns an element, a lane owns
// a 4-node line per phase; kernels_line3d.cuh) rebuilt around occupancy. ncu on k_line3d (profiles/
// r1_ncu_line3d_v5_l6.json) showed the FP64 pipe 50 % busy with 2 warps per scheduler: one warp's instruction stream
// needs ~5200 issue cycles per element pair of which ~2700 are FP64 issue, and two warps do not cover each other's
// shared-memory / MUFU / dependency waits. This kernel runs 3 warps per scheduler (12 per SM, 3 CTAs x 4 warps):
//  * <= 168 registers: the 8 pair fluxes of a line are evaluated as two batches of 4 -- batch A = {(lo,0), (0,1),
//    (0,2), (0,3)} completes node 0, batch B = {(1,2), (1,3), (2,3), (3,hi)} completes nodes 1..3 -- and the running
//    sums of earlier phases are added at the END of a phase (read-modify-write of the shared accumulator tile), so at
//    most 4 flux chains + 5 nodes + 3 partial nodes are live at a time;
//  * 9 KiB of shared memory per element instead of 11.5: phases run z, y, x. The z-line owner of the first phase is
//    also the conflict-free reader of the AoS block (stride-5 doubles), so the block is converted to the swizzled SoA
//    tile IN PLACE and its q never has to be re-read for the first phase; the x-line owner of the last phase holds 20
//    CONTIGUOUS doubles of du (nodes (0..3, j, k) x 5 variables), which go straight to global memory as 16-byte stores:
//    no output tile. The next element's block is prefetched into the q tile as soon as the last phase has read it;
//  * the swizzle is an XOR, p(i,j,k) = 16 k + 4 (j ^ k) + (i ^ k): conflict-free for x-, y- and z-line access like the
//    additive one, and the four positions of a line are P ^ (m * stride), stride = 1, 4, 21 -- three LOP3 per phase.
// Everything else (cp.async trace windows, rotated component slots so that one copy of the flux code serves all three
// directions, Taylor-branch means with the logarithmic branch as an out-of-line correction) is as in k_line3d.
// Reference stages covered: src/solvers/dg_3d_kernel.jl:188-257 (volume_flux_integral_kernel!), :1121-1152
// (prolong_interfaces_kernel!), :1155-1225 (surface_flux_kernel!, interface_flux_kernel!), :1773-1799
// (surface_integral_kernel!), :1802-1818 (jacobian_kernel!), :1821-1844 (source_terms_kernel!).
#pragma once
#include "kernels_line3d.cuh"


namespace tb {

constexpr int L6_TILE = 320;                          // 5 x 64 doubles
constexpr int L6_TR = L3_TR;                          // 512: x faces 2 x 96, y and z faces 4 x 80
constexpr int L6_EL = 2 * L6_TILE + L6_TR;            // 1152 doubles = 9 KiB per element
constexpr size_t l6_smem(int warps) { return (size_t)2 * L6_EL * warps * sizeof(double); }

struct L6Pair { double a[5], b[5]; };
// f_exact - f_taylor of one flagged pair (cold)
__device__ __noinline__ L3Vec5 l6_pair_correction(L6Pair p, double inv_gm1) {
  L3Vec5 r;
  l3_pair_correction(p.a, p.b, inv_gm1, r.v);
  return r;
}

// The 8 pairs of a line in evaluation order, as indices into the virtual nodes Q[0..5] = (low neighbour, nodes 0..3,
// high neighbour): all pairs of node 0 first (they complete node 0), then node 1's, then the rest.
__device__ constexpr int L6_PA[8] = {0, 1, 1, 1, 2, 2, 3, 4};
__device__ constexpr int L6_PB[8] = {1, 2, 3, 4, 3, 4, 4, 5};

// high word of 1e-4: f2 >= 1e-4 decided on the high words (both branches of the means are accurate near the threshold;
// NaN compares as "rough"; f2, g2 are squares, so the unsigned comparison is the ordering of the doubles)
constexpr unsigned L6_ROUGH_HI = 0x3F1A36E2u;

// NP two-point fluxes (orientation 1 in the rotated frame) of the pairs K0 .. K0 + NP - 1, staged across the pairs so
// that consecutive instructions are independent. hw[k] = larger high word of the two squared relative jumps of pair k:
// the pair needs the logarithmic branch iff hw[k] >= L6_ROUGH_HI (one integer max per pair on the hot path; the caller
// reduces them to ONE comparison per batch).
// SEEDS: the high words of the two reciprocal seeds of every pair were computed ahead (l6_seeds), the MUFU latency is
// then outside this function.
template <int NP> TB_D void l6_seeds(const double (&Q)[6][5], int (&seed_s)[NP], int (&seed_t)[NP]) {
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    const double* a = Q[L6_PA[k]]; const double* b = Q[L6_PB[k]];
    double r, rt;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a[0] + b[0]));
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rt) : "d"(a[0] * b[4] + b[0] * a[4]));
    seed_s[k] = __double2hiint(r); seed_t[k] = __double2hiint(rt);
  }
}
template <bool FAST, int K0, int NP, bool SEEDS = false>
TB_D void l6_fluxes(const double (&Q)[6][5], int vflux, int sflux, const EqPrm& prm, double (&F)[NP][5],
                    unsigned (&hw)[NP], const int* seed_s = nullptr, const int* seed_t = nullptr) {
  if (FAST) {
    double s[NP], r[NP], dd[NP], tt[NP], rt[NP], xy[NP], rm[NP], im[NP];
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      const double* a = Q[L6_PA[K0 + k]]; const double* b = Q[L6_PB[K0 + k]];
      s[k] = a[0] + b[0];
      dd[k] = a[0] - b[0];
      const double x = a[0] * b[4], y = b[0] * a[4];
      tt[k] = x + y;
      xy[k] = x - y;
      if (SEEDS) {
        r[k] = __hiloint2double(seed_s[K0 + k], 0);
        rt[k] = __hiloint2double(seed_t[K0 + k], 0);
      } else {
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r[k]) : "d"(s[k]));
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rt[k]) : "d"(tt[k]));
      }
    }
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      r[k] = fma(r[k], fma(-s[k], r[k], 1.0), r[k]);
      const double e1 = fma(-tt[k], rt[k], 1.0);
      rt[k] = fma(rt[k], fma(e1, e1, e1), rt[k]);
    }
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      const double uu = dd[k] * r[k], f2 = uu * uu;
      const double ut = xy[k] * rt[k], g2 = ut * ut;
      rm[k] = s[k] * fma(f2, fma(f2, fma(f2, -22.0 / 945, -2.0 / 45), -1.0 / 6), 0.5);
      im[k] = ((Q[L6_PA[K0 + k]][4] * Q[L6_PB[K0 + k]][4]) * rt[k]) * fma(g2, fma(g2, fma(g2, 2.0 / 7, 2.0 / 5), 2.0 / 3), 2.0);
      hw[k] = max((unsigned)__double2hiint(f2), (unsigned)__double2hiint(g2));
    }
#pragma unroll
    for (int k = 0; k < NP; ++k)
      l3_ranocha_from_means<1>(Q[L6_PA[K0 + k]], Q[L6_PB[K0 + k]], rm[k], im[k], prm.inv_gm1, F[k]);
  } else {
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      const bool surf = (K0 + k == 0) || (K0 + k == 7);
      l3_flux<1, -1>(surf ? sflux : vflux, Q[L6_PA[K0 + k]], Q[L6_PB[K0 + k]], prm, F[k]);
      hw[k] = 0;
    }
  }
}

// Fused low-storage Runge-Kutta stage (2N scheme, e.g. CarpenterKennedy2N54): instead of writing du the x phase does
//   tmp = a * tmp + dt * du;   u_out = u_in + b * tmp
// on the 20 contiguous doubles every lane holds, so du never exists in memory and the separate update sweep (read du,
// tmp, u; write tmp, u: 200 B/DOF) disappears. u_out must not alias u_in (neighbours still read u_in).
struct RkArgs {
  double* tmp;       // [nunknowns] the 2N register; not read when a == 0 (first stage)
  double a, b, dt;
};

// Halo exchange inside the kernel (multi-GPU, one process per GPU, peers' buffers mapped with CUDA IPC): every CTA
// first packs its share of the cut-face traces STRAIGHT INTO THE PEERS' receive buffers over NVLink (plain stores to
// peer memory), the last CTA to finish raises this rank's flag word in every peer's memory to the epoch of this rhs!,
// then all warps work through the interior elements; a warp looks at its own flag words (written by the peers) only
// when it reaches the first element pair that touches a cut face. One launch per rhs!, no NCCL kernel, no host-side
// event between the exchange and the cut elements. The receive buffers are double-buffered by epoch parity: a peer
// can only start writing epoch e + 2 after it has seen this rank's flag of epoch e + 1, which is raised by a kernel
// that runs after the one reading epoch e has finished (stream order).
// No deadlock: packing never waits, so every CTA of every rank raises its share before anybody spins.
struct P2PArgs {
  int npeers = 0;                              // 0: no exchange in this launch
  int first_cut_pair = 0;                      // element pairs >= this index may read halo traces
  unsigned long long epoch = 0;
  const int* peer_first = nullptr;             // [npeers + 1] first send slot of every peer (prefix sums)
  const int* peer_rank = nullptr;              // [npeers]
  double* const* peer_dst = nullptr;           // [npeers] where this rank's first slot for that peer lands (peer memory)
  unsigned long long* const* peer_flag = nullptr;   // [npeers] this rank's flag word in the peer's memory
  const unsigned long long* my_flags = nullptr;     // [nranks] flag words the peers write here
  unsigned int* done_counter = nullptr;        // CTAs of this launch that have finished packing (reset by the last)
};
TB_D unsigned long long l6_ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
TB_D void l6_st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// registers per thread that let CTAS CTAs of WARPS warps share the 64 K registers of an SM (allocation unit: 8 per thread)
constexpr int l6_maxnreg(int warps, int ctas) {
  const int r = (65536 / (32 * warps * ctas)) / 8 * 8;
  return r > 255 ? 255 : r;
}

// PP ("ping-pong"): one CTA of 3 warpgroups per SM (12 warps, launched at 168 registers). A phase is split into a
// register-light part (gather q / traces / running sums from shared memory, hand-over, stores, cp.async issue: <= L6_PP_LO
// registers) and the flux part (8 staged pair fluxes + accumulation: L6_PP_HI registers). Only ONE warpgroup is in its
// flux part at a time: it takes the token (named barrier), grows to L6_PP_HI registers with setmaxnreg.inc -- the other
// two warpgroups have shrunk to L6_PP_LO with setmaxnreg.dec -- and hands registers and token on when its 8 fluxes are
// accumulated. Every SM sub-partition then holds one warp that feeds the FP64 pipe at its 2-cycle cadence and two warps
// whose shared-memory / address / copy instructions issue in the gaps, instead of two warps that are in the same kind
// of section half of the time (profiles/r1_line6_notes.md: 0.68 eligible warps per cycle, FP64 pipe 52 % busy).
// L6_PP_TOKENS = 1: one warpgroup at a time in its flux part (232 registers there, 136 elsewhere: 232 + 2 * 136 = 504 =
// 3 * 168). L6_PP_TOKENS = 2: two at a time (192 + 192 + 120 = 504), in rotation order: warpgroup k may enter its n-th
// flux part when warpgroup (k + 1) % 3 has left its previous one.
#ifndef L6_PP_TOKENS
#define L6_PP_TOKENS 1
#endif
#ifndef L6_PP_HI
#if L6_PP_TOKENS == 2
#define L6_PP_HI 192
#define L6_PP_LO 120
#else
#define L6_PP_HI 232
#define L6_PP_LO 136
#endif
#endif
template <int REGS> TB_D void l6_reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(REGS)); }
template <int REGS> TB_D void l6_reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(REGS)); }
TB_D int l6_pp_bar(int wg, int round) { return 1 + 2 * wg + (round & 1); }    // ids 1..6
TB_D void l6_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
TB_D void l6_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
// the same, issued once `dep` has been computed (ptxas moves arithmetic across a volatile asm, a data dependence pins it)
TB_D void l6_bar_arrive_after(int id, int n, double dep) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n), "d"(dep) : "memory");
}
#ifndef L6_PP_ARRIVE_K
#define L6_PP_ARRIVE_K 5     // the token is passed on when pair K of the 8 has been accumulated (8: behind setmaxnreg.dec)
#endif

// UNR: the three phases as three copies of the code (direction, slot rotation and the first / last phase special
// cases are compile-time) instead of one copy with run-time selection: fewer executed instructions, 2.2x the hot
// loop's code size (TRIXIB200_LINE_SHAPE=16; bench.py times both and keeps the faster one).
template <int VFLUX, int SFLUX, bool SFV, int WARPS, int CTAS, int NP, bool RK = false, bool TOUT = false,
          bool PP = false, int UNR = 0>
__global__ void __launch_bounds__(32 * WARPS) __maxnreg__(l6_maxnreg(WARPS, CTAS))
k_line6(const __grid_constant__ Dev d, const __grid_constant__ LineOps ops, double* __restrict__ du,
        const double* __restrict__ u, double t, const int* __restrict__ elems, int64_t count,
        const __grid_constant__ RkArgs rk, const __grid_constant__ P2PArgs p2p) {
  // TILE: the x phase parks its 20 doubles per lane in a padded shared tile and a second pass moves whole 16-byte
  // chunks l16 + 16 m, i.e. fully coalesced global accesses. A 16-byte access per lane at a 160-byte stride touches
  // ~40 cache lines per warp instruction and makes the L1 the bottleneck of the fused RK stage (7.6 vs 5.8 ms at
  // level 7); for the plain du store the tile is worth 1 % (profiles/r1_line6_notes.md).
  constexpr bool TILE = RK || TOUT;
  constexpr int NV = 5, NN = 64;
  constexpr bool FAST = (VFLUX == TRIXIB200_FLUX_RANOCHA && SFLUX == TRIXIB200_FLUX_RANOCHA);
  extern __shared__ __align__(16) double smem_l6[];
  // Per-lane constants live in two OPAQUE registers (a self-shuffle hides their origin): under register pressure ptxas
  // otherwise re-derives them from S2R SR_TID.X + an integer chain in every phase (short-scoreboard stalls at the top
  // of each phase in profiles/r1_ncu_line6_12warps_l6_hot.txt); from the packed word they cost one or two ALU ops.
  unsigned lp = threadIdx.x & 31;          // l16 | half << 4
  unsigned sb = (unsigned)__cvta_generic_to_shared(smem_l6) +
                (unsigned)((2 * (threadIdx.x >> 5) + ((threadIdx.x >> 4) & 1)) * L6_EL * sizeof(double));
  // (a self-shuffle: ptxas cannot see through it, an empty asm only stops the front end)
  lp = __shfl_sync(0xffffffffu, lp, threadIdx.x & 31);
  sb = __shfl_sync(0xffffffffu, sb, threadIdx.x & 31);
  // Everything derived from the lane id. PP re-derives it from the two opaque registers at the top of every phase and
  // again behind the flux part: ptxas otherwise hoists ~25 lane-dependent addresses out of the persistent loop and, as
  // the register-light parts have 136 registers, keeps them in LOCAL memory (one L2 round trip per reload: the first
  // PP build spent 25 % of its stall samples on them, profiles/r2_line6_pingpong.md).
  int half, l16, la, lb, Px, Py, Pz;
  double *sq, *sacc, *tr;
  auto relane = [&]() {
    unsigned a = lp, b = sb;
    if (PP) asm volatile("" : "+r"(a), "+r"(b));
    half = a >> 4; l16 = a & 15;
    la = l16 & 3; lb = l16 >> 2;
    double* ebase = reinterpret_cast<double*>(__cvta_shared_to_generic((size_t)b));
    sq = ebase;                    // AoS landing zone of the element block, then the swizzled SoA q [5][64]
    sacc = ebase + L6_TILE;        // swizzled SoA running sums [5][64], handed from phase to phase
    tr = ebase + 2 * L6_TILE;      // face traces of the six neighbours
    // tile positions of this lane's line nodes: P ^ (m * stride) with stride 1 (x), 4 (y), 21 (z)
    Px = 16 * lb + 4 * (la ^ lb) + lb; Py = 20 * lb + (la ^ lb); Pz = l16;
  };
  relane();
  const EqPrm prm = d.prm;
  const double gm1 = prm.gamma - 1;
  const int vflux = (VFLUX >= 0) ? VFLUX : d.vol_flux;
  const int sflux = (SFLUX >= 0) ? SFLUX : d.surf_flux;
  const int npairs = (int)((count + 1) >> 1);          // element ids are int (face_nbr), so pair indices fit
  const int wid = blockIdx.x * WARPS + (threadIdx.x >> 5), nw = gridDim.x * WARPS;

  auto elem_of = [&](int pr, bool& valid) -> int {
    int s = 2 * pr + half;
    valid = s < (int)count;
    if (!valid) s = (int)count - 1;
    return elems ? elems[s] : s;
  };
  // neighbour codes of the two faces of direction dir (24 B per element, 8 B aligned)
  auto load_codes = [&](int el, int dir) -> int2 {
    return reinterpret_cast<const int2*>(d.face_nbr + (size_t)el * 6)[dir];
  };
  // what the phases need to know about the faces of the CURRENT element: 4 bits per direction
  // (low face local, low face given flux, high face local, high face given flux)
  auto face_bits = [](int2 c, int dir) -> unsigned {
    // without given-flux faces only "the x faces are local" is ever read (the 48-byte window layout of their traces)
    if (UNR != 0 && !SFV) return dir == 0 ? ((c.x >= 0 ? 1u : 0u) | (c.y >= 0 ? 4u : 0u)) : 0u;
    return ((c.x >= 0 ? 1u : 0u) | (c.x == NB_SFV ? 2u : 0u) | (c.y >= 0 ? 4u : 0u) | (c.y == NB_SFV ? 8u : 0u)) << (4 * dir);
  };
  auto issue_block = [&](int e) {
    const double* ue = u + (size_t)NV * NN * e;
#pragma unroll
    for (int m = 0; m < 10; ++m) cp_async16(sq + 2 * (l16 + 16 * m), ue + 2 * (l16 + 16 * m));
  };
  auto face_off = [](int dir, int sd) { return dir == 0 ? 96 * sd : 32 + 160 * dir + 80 * sd; };
  // trace windows exactly as in k_line3d (x faces 16 windows of 48 B, y faces four runs of 160 B, z faces one run of
  // 640 B of the neighbour's block; given fluxes and halo traces are dense [f][v] runs of 640 B)
  auto issue_traces = [&](auto dir_tag, int e, const int c0, const int c1) {
    constexpr int dir = decltype(dir_tag)::value;
#pragma unroll
    for (int sd = 0; sd < 2; ++sd) {
      const int code = sd == 0 ? c0 : c1;
      double* dst = tr + face_off(dir, sd);
      const double* src;
      if (code >= 0) src = u + (size_t)NV * NN * code;
      else if (SFV && code == NB_SFV) src = d.sfv + (size_t)NV * 16 * (2 * dir + sd + (size_t)6 * e);
      else src = d.halo_recv + (size_t)nb_halo_slot(code) * 16 * NV;
      const bool local = code >= 0;
#pragma unroll
      for (int it = 0; it < 3; ++it) {
        const int c = l16 + 16 * it;
        int so = 2 * c, dof = 2 * c;
        bool on = c < 40;
        if (local) {
          if (dir == 0) {
            const int f = c / 3, w = c - 3 * f;
            so = 20 * f + (sd == 0 ? 14 : 0) + 2 * w; dof = 6 * f + 2 * w; on = true;
          } else if (dir == 1) {
            const int k = c / 10;
            so = 60 * k + (sd == 0 ? 60 : 0) + 2 * c;
          } else {
            so = (sd == 0 ? 240 : 0) + 2 * c;
          }
        }
        if (on) cp_async16(dst + dof, src + so);
      }
    }
  };
  using D0 = std::integral_constant<int, 0>;
  using D1 = std::integral_constant<int, 1>;
  using D2 = std::integral_constant<int, 2>;

  // ---- multi-GPU: pack the cut-face traces into the peers' receive buffers, then raise the flags (see P2PArgs)
  bool halo_ready = p2p.npeers == 0;
  if (p2p.npeers != 0) {
    const int lane = threadIdx.x & 31;
    const int nslots = p2p.peer_first[p2p.npeers];
    for (int sl = wid; sl < nslots; sl += nw) {
      int k = 0;
      while (sl >= p2p.peer_first[k + 1]) ++k;
      double* dst = p2p.peer_dst[k] + (size_t)(sl - p2p.peer_first[k]) * (16 * NV);
      const int se = d.send_elem[sl], sdir = d.send_dir[sl];
      const double* ue = u + (size_t)NV * NN * se;
      const int fixed = (sdir & 1) ? 3 : 0, sdim = sdir >> 1;
#pragma unroll
      for (int it2 = 0; it2 < 3; ++it2) {
        const int i = lane + 32 * it2;
        if (i < 16 * NV) {
          const int f = i / NV, v = i - NV * f;
          dst[i] = ue[NV * face_node<3>(4, sdim, fixed, f) + v];      // layout [f][v] like k_pack_halo
        }
      }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned old = atomicAdd(p2p.done_counter, 1u);
      if (old == gridDim.x - 1) {                      // every CTA of this launch has packed and fenced
        *p2p.done_counter = 0;                         // the next launch is ordered behind this one
        __threadfence_system();
        for (int k = 0; k < p2p.npeers; ++k) l6_st_release_sys(p2p.peer_flag[k], p2p.epoch);
      }
    }
  }
  // a warp calls this before it touches the first halo trace (prefetch of a pair >= first_cut_pair)
  auto wait_halo = [&]() {
    if (!halo_ready) {
      for (int k = (int)(threadIdx.x & 31); k < p2p.npeers; k += 32)
        while (l6_ld_acquire_sys(p2p.my_flags + p2p.peer_rank[k]) < p2p.epoch) {}
      __syncwarp();
      halo_ready = true;
    }
  };
  static_assert(!PP || (WARPS == 12 && CTAS == 1 && NP >= 4), "ping-pong shape: 3 warpgroups, one CTA per SM");
  // PP: every warp of the CTA runs the same number of iterations (the token goes round the three warpgroups in a fixed
  // order); warps whose pair index runs past the end repeat the last element with their stores switched off
  const int wg = (threadIdx.x >> 7);
  const int niter = PP ? (npairs + nw - 1) / nw : 0;
  if (PP) {
    l6_reg_dec<L6_PP_LO>();
    // barrier l6_pp_bar(k, n) lets warpgroup k into its n-th flux part; it is raised by warpgroup (k + 3 - TOKENS) % 3
    // when that one leaves a flux part, and once at the start for the first TOKENS warpgroups. With two tokens a
    // warpgroup can be handed its next TWO entries before it takes the first (never three), so consecutive entries
    // use different barriers (parity of n): a named barrier must not collect two rounds of arrivals.
    for (int k = 0; k < L6_PP_TOKENS; ++k)
      if (wg == (k + 3 - L6_PP_TOKENS) % 3) l6_bar_arrive(l6_pp_bar(k, 0), 256);
  }
  unsigned fbits = 0;
  int pr = wid;
  bool valid = false;
  int e = 0;
  // cp.async groups are committed in the order  z-traces, y-traces, block, x-traces  (of the NEXT pair): at the top
  // of an iteration only the x-traces may be pending (wait_group 1), at the x phase the z- and y-traces of the next
  // pair (wait_group 2)
  if (PP || pr < npairs) {
    if (pr >= p2p.first_cut_pair) wait_halo();
    e = elem_of(pr, valid);
    const int2 cx = load_codes(e, 0), cy = load_codes(e, 1), cz = load_codes(e, 2);
    fbits = face_bits(cx, 0) | face_bits(cy, 1) | face_bits(cz, 2);
    issue_traces(D2{}, e, cz.x, cz.y); cp_async_commit();
    issue_traces(D1{}, e, cy.x, cy.y); cp_async_commit();
    issue_block(e); cp_async_commit();
    issue_traces(D0{}, e, cx.x, cx.y); cp_async_commit();
  }

  for (int it = 0; PP ? it < niter : pr < npairs; pr += nw, ++it) {
    // the last iteration prefetches its own element again (PP: elem_of clamps a pair index past the end)
    const int pr_next = (PP || pr + nw < npairs) ? pr + nw : pr;
    bool valid_next = false;
    const int e_next = elem_of(pr_next, valid_next);
    unsigned fbits_next = 0;
    const double inv_jac = d.inv_jac[e];
    if (PP) relane();

    // ---- the block has landed: cons -> q in place. Lane (la, lb) reads the nodes (la, lb, m) = its z-line.
    cp_async_wait<1>();
    __syncwarp();
    double Q[6][NV];
    {
      double un[4][NV];
#pragma unroll
      for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int v = 0; v < NV; ++v) un[m][v] = sq[NV * (l16 + 16 * m) + v];
      __syncwarp();
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        double q[NV];
        l3_to_q(un[m], gm1, q);
        const int pos = Pz ^ (21 * m);
#pragma unroll
        for (int v = 0; v < NV; ++v) sq[v * NN + pos] = q[v];
        if (true) {   // z phase first: slot 1 = v3/2, then (v1/2, v2/2)
          Q[1 + m][0] = q[0]; Q[1 + m][1] = q[3]; Q[1 + m][2] = q[1]; Q[1 + m][3] = q[2]; Q[1 + m][4] = q[4];
        }
      }
    }
    __syncwarp();

    // UNR = 0: one copy of the phase code, run-time direction. UNR = 1: three copies. UNR = 2: the z phase (no running
    // sums to load, q still in registers) as its own copy, the y and x phases share the second one.
    if constexpr (UNR == 2) {
      {
        constexpr int step = 0;
#include "kernels_line6_phase.inc"
      }
#pragma unroll 1
      for (int step = 1; step < 3; ++step) {
#include "kernels_line6_phase.inc"
      }
    } else {
#pragma unroll(UNR == 1 ? 3 : 1)
      for (int step = 0; step < 3; ++step) {
#include "kernels_line6_phase.inc"
      }
    }
    e = e_next; valid = valid_next; fbits = fbits_next;
  }
  cp_async_wait<0>();
  if (PP && wg < L6_PP_TOKENS) l6_bar_sync(l6_pp_bar(wg, 3 * niter), 256);   // the token passed on after the CTA's last flux parts
}

// ---------------------------------------------------------------------------------------------- host side
// Launch shapes. Measured on B200 at level 7 (profiles/r1_line6_shapes.txt): 2 CTAs x 4 warps with all 8 pairs staged
// (<= 255 registers) 4.75 ms, batches of 4 / 2 pairs 4.87 / 4.89 ms; 3 CTAs x 4 warps (168 registers) 4.87 ms with
// batches of 2 and 5.3 ms with batches of 4 (local-memory spills of loop state, each reload an L2 round trip). The
// kernel's time follows the SUM of the issue costs of its instructions, not the occupancy: 1 warp per scheduler already
// reaches 70 % of the throughput of 2, and 3 add nothing (profiles/r1_line6_notes.md).
template <int VFLUX, int SFLUX, bool SFV, int CTAS, int WARPS, int NP, bool RK = false, bool TOUT = false,
          bool PP = false, int UNR = 0>
static int line6_launch_t(const Dev& d, const LineOps& ops, double* du, const double* u, double t, const int* elems,
                          int64_t count, cudaStream_t stream, int sm_count, const RkArgs& rk = RkArgs{nullptr, 0, 0, 0},
                          const P2PArgs& p2p = P2PArgs{}) {
  auto kern = k_line6<VFLUX, SFLUX, SFV, WARPS, CTAS, NP, RK, TOUT, PP, UNR>;
  static DeviceOnce configured;
  if (configured.need()) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l6_smem(WARPS)) != cudaSuccess)
      return TRIXIB200_ECUDA;
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  }
  if (count <= 0) return 0;
  const int64_t npairs = (count + 1) / 2;
  const int64_t want = (npairs + WARPS - 1) / WARPS;
  const unsigned blocks = (unsigned)std::min<int64_t>(want, (int64_t)sm_count * CTAS);
  kern<<<blocks, 32 * WARPS, l6_smem(WARPS), stream>>>(d, ops, du, u, t, elems, count, rk, p2p);
  return cudaGetLastError() == cudaSuccess ? 0 : TRIXIB200_ECUDA;
}

static int line6_launch(const trixib200_config& c, const Dev& d, const LineOps& ops, double* du, const double* u,
                        double t, const int* elems, int64_t count, cudaStream_t s, int sm_count,
                        const P2PArgs& p2p = P2PArgs{}) {
  const RkArgs rk{nullptr, 0, 0, 0};
  constexpr int R = TRIXIB200_FLUX_RANOCHA;
  const bool sfv = d.B > 0 || d.M > 0;   // faces whose flux is given in surface_flux_values (boundaries, mortars)
  if (c.volume_flux == R && c.surface_flux == R) {
    // TRIXIB200_LINE_SHAPE=3: 3 CTAs x 4 warps per SM at 168 registers, batches of 2 pairs (A/B measurements)
    static const bool three = getenv("TRIXIB200_LINE_SHAPE") && atoi(getenv("TRIXIB200_LINE_SHAPE")) == 3;
    if (three && !sfv) return line6_launch_t<R, R, false, 3, 4, 2>(d, ops, du, u, t, elems, count, s, sm_count, rk, p2p);
    // TRIXIB200_LINE_SHAPE=8: du stored directly from the x-line owners (160-byte runs per lane) instead of through
    // the shared tile (A/B measurements: 4.71 vs 4.66 ms at level 7)
    static const bool direct = getenv("TRIXIB200_LINE_SHAPE") && atoi(getenv("TRIXIB200_LINE_SHAPE")) == 8;
    if (direct && !sfv) return line6_launch_t<R, R, false, 2, 4, 8>(d, ops, du, u, t, elems, count, s, sm_count, rk, p2p);
    // TRIXIB200_LINE_SHAPE=12: the ping-pong shape (3 warpgroups, flux parts serialised by a token, setmaxnreg)
    static const bool pp = getenv("TRIXIB200_LINE_SHAPE") && atoi(getenv("TRIXIB200_LINE_SHAPE")) == 12;
    if (pp && !sfv) return line6_launch_t<R, R, false, 1, 12, 8, false, true, true>(d, ops, du, u, t, elems, count, s, sm_count, rk, p2p);
    // TRIXIB200_LINE_SHAPE=16: the default shape with the three phases unrolled
    static const bool unr = getenv("TRIXIB200_LINE_SHAPE") && atoi(getenv("TRIXIB200_LINE_SHAPE")) == 16;
    if (unr && !sfv) return line6_launch_t<R, R, false, 2, 4, 8, false, true, false, 1>(d, ops, du, u, t, elems, count, s, sm_count, rk, p2p);
    // TRIXIB200_LINE_SHAPE=17: the z phase peeled, the y and x phases in one copy
    static const bool peel = getenv("TRIXIB200_LINE_SHAPE") && atoi(getenv("TRIXIB200_LINE_SHAPE")) == 17;
    if (peel && !sfv) return line6_launch_t<R, R, false, 2, 4, 8, false, true, false, 2>(d, ops, du, u, t, elems, count, s, sm_count, rk, p2p);
    return sfv ? line6_launch_t<R, R, true, 2, 4, 8, false, true>(d, ops, du, u, t, elems, count, s, sm_count)
               : line6_launch_t<R, R, false, 2, 4, 8, false, true>(d, ops, du, u, t, elems, count, s, sm_count, rk, p2p);
  }
  return line6_launch_t<-1, -1, true, 2, 4, 2, false, true>(d, ops, du, u, t, elems, count, s, sm_count, rk, p2p);
}

// rhs! fused with the 2N Runge-Kutta stage update: u_out = u_in + b * (tmp = a * tmp + dt * rhs(u_in, t))
static int line6_launch_rk(const trixib200_config& c, const Dev& d, const LineOps& ops, double* u_out, const double* u_in,
                           double t, const int* elems, int64_t count, cudaStream_t s, int sm_count, const RkArgs& rk,
                           const P2PArgs& p2p = P2PArgs{}) {
  constexpr int R = TRIXIB200_FLUX_RANOCHA;
  const bool sfv = d.B > 0 || d.M > 0;
  if (c.volume_flux == R && c.surface_flux == R)
    return sfv ? line6_launch_t<R, R, true, 2, 4, 8, true>(d, ops, u_out, u_in, t, elems, count, s, sm_count, rk)
               : line6_launch_t<R, R, false, 2, 4, 8, true>(d, ops, u_out, u_in, t, elems, count, s, sm_count, rk, p2p);
  return line6_launch_t<-1, -1, true, 2, 4, 2, true>(d, ops, u_out, u_in, t, elems, count, s, sm_count, rk, p2p);
}

}  // namespace tb
