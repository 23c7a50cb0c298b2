// Multi-stream H-solve for the shipped geometry (F = 512 + E rows, E <= 1; R_x = 100 speech atoms, R_d = 100 noise atoms of
// which the first R_a = 50 are adapted per stream).
//
// Of the 200 columns of W = [B_x B_d] (sparse_nmf.m:186-208 with bnmf_sep_event_RT_IS16.m:150-156) only the R_a adapted
// noise atoms differ between streams: B_x and B_d(:, R_a+1:end) are the same for every stream of a batch.  One 8-CTA
// cluster therefore iterates S streams in lock step:
//   * CTA c owns rows [64c, 64c+64).  Its slice of the 150 SHARED columns lives in REGISTERS, twice, as FP64 tensor-core
//     fragments (mma.sync.m8n8k4.f64, N = streams): once in the A layout of  Lambda = W h  (M = rows, K = atoms) and once in
//     the A layout of  g = W'(v./Lambda)  (M = atoms, K = rows).  No shared-memory traffic for 3/4 of the flops.
//   * the PRIVATE columns (S x 50 x 64 doubles) sit in shared memory with the row-pair XOR swizzle of hsolve_fast_kernel and
//     are consumed as mat-vecs (lanes <-> row pairs for Lambda, lanes <-> atoms for g), un-normalised: the column scaling of
//     sparse_nmf.m:157-160 is folded into the small vectors (h / wn on the way in, g / wn on the way out).
//   * CTA c is also the OWNER of stream c: per iteration every CTA sends its partial g of stream c (204 doubles) to CTA c
//     with ONE bulk copy over distributed shared memory (cp.async.bulk, bytes counted on the owner's mbarrier); the owner adds
//     the 8 partials in rank order, evaluates the cost / stop rule (sparse_nmf.m:250-283), updates h (:195) and sends
//     h ./ wn + its done flag back to all 8 CTAs with one bulk copy each.  Streams that stopped keep their h; the cluster
//     leaves the loop when all S are done, then one more pass gives B_x A_x and B_d A_d (bnmf_sep_event_RT_IS16.m:174,197).
// Results per stream do not depend on which streams share a cluster.
#include <cooperative_groups.h>
#include <cmath>
#include "online.cuh"
#include "online_dev.cuh"

namespace cg = cooperative_groups;

namespace snmfnat {

constexpr int MS_CL = 8;         // CTAs per cluster
constexpr int MS_ROWS = 64;      // rows per CTA
constexpr int MS_THREADS = 256;
constexpr int MS_WARPS = 8;
constexpr int MS_RX = 100, MS_RF = 50, MS_RA = 50;
constexpr int MS_RS = MS_RX + MS_RF;    // shared atoms
constexpr int MS_KS = 152;              // ... padded to whole 8-atom tiles
constexpr int MS_KT = MS_KS / 8;        // 19 atom tiles
constexpr int MS_HLD = 216;             // row stride of hS / gpart (== 8 mod 16: conflict-free 16-byte fragment loads)
constexpr int MS_ROWLEN = 204;          // doubles per exchanged row: [0,152) shared atoms, [152,202) private, [202] flag / cost
constexpr int MS_PRIV0 = 152, MS_FLAG = 202;
constexpr unsigned MS_ROWBYTES = MS_ROWLEN * 8;
constexpr int MS_RLD = 72;              // row stride of rS (== 8 mod 16)
constexpr int MS_LLD = 66;              // row stride of lam_p (== 2 mod 8)
constexpr int MS_XB = MS_RX / 8, MS_XR = MS_RX % 8;   // the x | d boundary inside atom tile XB
static_assert(MS_XR % 2 == 0, "the two atoms of a lane's pair must belong to the same class");

template <int S>
struct MsLayout {
  static constexpr size_t off_Wp = 0;                                          // [S][RA][64] swizzled private columns
  static constexpr size_t off_hS = off_Wp + (size_t)S * MS_RA * MS_ROWS;       // [8][HLD]  h./wn per stream; aliased by gpart
  static constexpr size_t off_recv = off_hS + 8 * MS_HLD;                      // [8][ROWLEN] partials of the owned stream
  static constexpr size_t off_stage = off_recv + 8 * MS_ROWLEN;                // [ROWLEN] the owner's outgoing row
  static constexpr size_t off_lamp = off_stage + MS_ROWLEN;                    // [8][LLD] private part of Lambda
  static constexpr size_t off_rS = off_lamp + 8 * MS_LLD;                      // [8][RLD] ratio v./Lambda
  static constexpr size_t off_costw = off_rS + 8 * MS_RLD;                     // [9][8] cost partials per warp (+ tail row)
  static constexpr size_t off_WnS = off_costw + 72;                            // [KS] tail row of the shared columns
  static constexpr size_t off_WnP = off_WnS + MS_KS;                           // [S][RA] tail row of the private columns
  static constexpr size_t off_rN = off_WnP + (size_t)((S * MS_RA + 1) & ~1);   // [8] ratio of the tail row, [8] Lambda of it
  static constexpr size_t off_hsum = off_rN + 16;                              // [2][8]
  static constexpr size_t off_bar = off_hsum + 16;                             // 2 mbarriers
  static constexpr size_t off_slot = off_bar + 2;                              // 8 ints
  static constexpr size_t doubles = off_slot + 4;
  static constexpr size_t bytes = doubles * sizeof(double);
};

// column norms and sums of the shared columns: colstat[a] = ||W(:,a)||, colstat[KS + a] = sum(W(:,a))
__global__ void ms_colstat_kernel(const double* __restrict__ Bx, const double* __restrict__ Bd_fix, int F, int LDF,
                                  double* __restrict__ colstat) {
  const int a = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (a >= MS_KS) return;
  double s1 = 0.0, s2 = 0.0;
  if (a < MS_RS) {
    const double* src = a < MS_RX ? Bx + (size_t)a * LDF : Bd_fix + (size_t)(MS_RA + a - MS_RX) * LDF;
    for (int f = lane; f < F; f += 32) {
      const double x = src[f];
      s1 += x;
      s2 = fma(x, x, s2);
    }
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if (lane == 0) {
    colstat[a] = a < MS_RS ? sqrt(s2) : 1.0;
    colstat[MS_KS + a] = s1;
  }
}

// MODE 0: all atoms; 1: speech atoms only (B_x A_x); 2: noise atoms only (B_d A_d)
template <int MODE>
__device__ __forceinline__ void ms_pass_a_shared(const double (&Wa)[2 * MS_KT], const double* __restrict__ hb, int lj,
                                                 double& c0, double& c1) {
  double p0 = 0.0, p1 = 0.0, q0 = 0.0, q1 = 0.0;
#pragma unroll
  for (int u = 0; u < MS_KT; ++u) {
    if (MODE == 1 && (u > MS_XB || (u == MS_XB && MS_XR == 0))) continue;
    if (MODE == 2 && u < MS_XB) continue;
    double2 hh = *reinterpret_cast<const double2*>(hb + 8 * u);
    if (MODE != 0 && u == MS_XB) {
      const bool is_x = 2 * lj < MS_XR;
      if ((MODE == 1) != is_x) hh = make_double2(0.0, 0.0);
    }
    if (u & 1) {
      dmma884(q0, q1, Wa[2 * u], hh.x);
      dmma884(q0, q1, Wa[2 * u + 1], hh.y);
    } else {
      dmma884(p0, p1, Wa[2 * u], hh.x);
      dmma884(p0, p1, Wa[2 * u + 1], hh.y);
    }
  }
  c0 = p0 + q0;
  c1 = p1 + q1;
}

// private part of Lambda for stream `warp`: lanes <-> row pairs; result to lam_p[warp][2*lane .. +1]
__device__ __forceinline__ void ms_pass_a_private(const double* __restrict__ wp, const double* __restrict__ hp, int lane,
                                                  double* __restrict__ out) {
  double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
#pragma unroll 5
  for (int k = 0; k < MS_RA; k += 2) {
    const double2 hh = *reinterpret_cast<const double2*>(hp + k);
    const double2 w0 = *reinterpret_cast<const double2*>(wp + (size_t)k * MS_ROWS + 2 * (lane ^ (k & 7)));
    const double2 w1 = *reinterpret_cast<const double2*>(wp + (size_t)(k + 1) * MS_ROWS + 2 * (lane ^ ((k + 1) & 7)));
    a0 = fma(w0.x, hh.x, a0);
    a1 = fma(w0.y, hh.x, a1);
    b0 = fma(w1.x, hh.y, b0);
    b1 = fma(w1.y, hh.y, b1);
  }
  *reinterpret_cast<double2*>(out + 2 * lane) = make_double2(a0 + b0, a1 + b1);
}

template <int NT>
__device__ __forceinline__ void ms_pass_b_shared(const double (&Wb)[3][16], const double* __restrict__ rb,
                                                 double (&g)[3][2]) {
#pragma unroll
  for (int q = 0; q < 3; ++q) g[q][0] = g[q][1] = 0.0;
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const double2 rr = *reinterpret_cast<const double2*>(rb + 8 * u);
#pragma unroll
    for (int q = 0; q < NT; ++q) dmma884(g[q][0], g[q][1], Wb[q][2 * u], rr.x);
#pragma unroll
    for (int q = 0; q < NT; ++q) dmma884(g[q][0], g[q][1], Wb[q][2 * u + 1], rr.y);
  }
}

template <int S>
__global__ void __cluster_dims__(MS_CL, 1, 1) __launch_bounds__(MS_THREADS, 1)
hsolve_ms_kernel(OnlineDims d, OnlineScalars sc, SlotState st, FrameArrays fr, const double* __restrict__ h_init, int g_step,
                 const double2* __restrict__ log_tab, const double* __restrict__ colstat, const int* __restrict__ perm,
                 int n_active) {
  using L = MsLayout<S>;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int grp = (int)(blockIdx.x / MS_CL);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int li = lane >> 2, lj = lane & 3;
  const int LDF = d.LDF, R = d.R;
  const int f0 = rank * MS_ROWS;
  const bool tail_rank = (rank == MS_CL - 1) && d.F > MS_CL * MS_ROWS;   // the CTA that also carries row 512
  const int FT = MS_CL * MS_ROWS;                                        // index of the tail row
  const double flr = sc.flr;

  extern __shared__ __align__(1024) double smem[];
  double* Wp = smem + L::off_Wp;
  double* hS = smem + L::off_hS;
  double* gpart = hS;                   // same rows: a row is h./wn from the owner's send until pass B overwrites it
  double* recv = smem + L::off_recv;
  double* stage = smem + L::off_stage;
  double* lam_p = smem + L::off_lamp;
  double* rS = smem + L::off_rS;
  double* costw = smem + L::off_costw;
  double* WnS = smem + L::off_WnS;
  double* WnP = smem + L::off_WnP;
  double* rN = smem + L::off_rN;
  double* lamN = rN + 8;
  double* hsumw = smem + L::off_hsum;
  int* slot_s = reinterpret_cast<int*>(smem + L::off_slot);
  const unsigned barRS = (unsigned)__cvta_generic_to_shared(smem + L::off_bar);
  const unsigned barAG = barRS + 8u;

  // ---- the streams of this cluster ----
  if (tid < 8) {
    int slot = -1;
    const int p = grp * S + tid;
    if (tid < S && p < n_active) {
      const int idx = perm ? perm[p] : p;
      slot = d.slot0 + idx * d.slot_stride;
      const int l = g_step + 1 - st.l_offset[slot];
      if (l < 1 || l > st.n_hops[slot]) slot = -1;
    }
    slot_s[tid] = slot;
  }
  __syncthreads();
  bool any = false;
#pragma unroll
  for (int n = 0; n < S; ++n) any |= slot_s[n] >= 0;
  if (!any) return;  // uniform over the cluster
  const int my_slot = rank < S ? slot_s[rank] : -1;   // the stream this CTA owns

  if (tid == 0) {
    hf_mbar_init(barRS, 1);
    hf_mbar_init(barAG, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  // ---- private columns: 16-byte cp.async into the swizzled layout (pair p of atom k at slot p ^ (k & 7)) ----
  for (int c = tid; c < S * MS_RA * 32; c += MS_THREADS) {
    const int n = c / (MS_RA * 32), rem = c - n * (MS_RA * 32);
    const int k = rem >> 5, p = rem & 31;
    const int slot = slot_s[n];
    double* dst = Wp + ((size_t)(n * MS_RA + k) * 32 + (p ^ (k & 7))) * 2;
    if (slot >= 0) {
      const double* src = st.Bd[st.bd_sel[slot]] + ((size_t)slot * d.R_d + k) * LDF + f0 + 2 * p;
      const unsigned da = (unsigned)__cvta_generic_to_shared(dst);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(da), "l"(src) : "memory");
    } else {
      *reinterpret_cast<double2*>(dst) = make_double2(0.0, 0.0);
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");

  // ---- shared columns -> registers, in both fragment layouts ----
  auto shared_col = [&](int a) -> const double* {
    return a < MS_RX ? st.Bx + (size_t)a * LDF : st.Bd_fix + (size_t)(MS_RA + a - MS_RX) * LDF;
  };
  double Wa[2 * MS_KT];   // Wa[2u+e] = W[f0 + 8 warp + li][8u + 2 lj + e]
#pragma unroll
  for (int u = 0; u < MS_KT; ++u)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int a = 8 * u + 2 * lj + e;
      Wa[2 * u + e] = a < MS_RS ? shared_col(a)[f0 + 8 * warp + li] : 0.0;
    }
  double Wb[3][16];       // Wb[q][2u+e] = W[f0 + 8u + 2 lj + e][8 (warp + 8q) + li]
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const int a = 8 * (warp + 8 * q) + li;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      double2 w = make_double2(0.0, 0.0);
      if (warp + 8 * q < MS_KT && a < MS_RS) w = *reinterpret_cast<const double2*>(shared_col(a) + f0 + 8 * u + 2 * lj);
      Wb[q][2 * u] = w.x;
      Wb[q][2 * u + 1] = w.y;
    }
  }
  // V of the lane's two (row, stream) elements of the accumulator tile
  double v[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int n = 2 * lj + e;
    const int slot = n < S ? slot_s[n] : -1;
    v[e] = 1.0;
    if (slot >= 0) v[e] = fmax(fr.Ym[(size_t)(st.frame_base[slot] + g_step) * LDF + f0 + 8 * warp + li], flr);  // sparse_nmf.m:169
  }
  double vN = 1.0;        // tail row: lane n of warp 7
  if (tail_rank && warp == MS_WARPS - 1 && lane < S && slot_s[lane] >= 0)
    vN = fmax(fr.Ym[(size_t)(st.frame_base[slot_s[lane]] + g_step) * LDF + FT], flr);

  // ---- small buffers ----
  for (int i = tid; i < 8 * MS_HLD; i += MS_THREADS) hS[i] = 0.0;
  for (int i = tid; i < 8 * MS_RLD; i += MS_THREADS) rS[i] = 0.0;
  for (int i = tid; i < 8 * MS_LLD; i += MS_THREADS) lam_p[i] = 0.0;
  if (tid < 72) costw[tid] = 0.0;
  if (tid < 16) {
    rN[tid] = 0.0;
    hsumw[tid] = 0.0;
  }
  for (int a = tid; a < MS_KS; a += MS_THREADS) WnS[a] = (tail_rank && a < MS_RS) ? shared_col(a)[FT] : 0.0;
  for (int i = tid; i < S * MS_RA; i += MS_THREADS) {
    const int n = i / MS_RA, k = i - n * MS_RA;
    const int slot = slot_s[n];
    WnP[i] = (tail_rank && slot >= 0) ? st.Bd[st.bd_sel[slot]][((size_t)slot * d.R_d + k) * LDF + FT] : 0.0;
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  cluster.sync();   // barriers initialised everywhere, local buffers staged

  // ---- exchange helpers ----
  unsigned parRS = 0, parAG = 0;
  // every CTA sends its partial row of stream o to the owner o; the owner's barrier counts 8 rows
  auto push_partials = [&]() {
    hf_fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      if (rank < S) hf_mbar_expect_tx(barRS, 8u * MS_ROWBYTES);
      hf_mbar_expect_tx(barAG, (unsigned)S * MS_ROWBYTES);
      const unsigned dst = (unsigned)__cvta_generic_to_shared(recv + (size_t)rank * MS_ROWLEN);
#pragma unroll
      for (int o = 0; o < S; ++o)
        hf_bulk_push(hf_mapa(dst, o), (unsigned)__cvta_generic_to_shared(gpart + (size_t)o * MS_HLD), MS_ROWBYTES,
                     hf_mapa(barRS, o));
    }
  };
  // the owner sends its staged row to row `rank` of hS in all 8 CTAs
  auto push_row = [&]() {
    hf_fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      const unsigned dst = (unsigned)__cvta_generic_to_shared(hS + (size_t)rank * MS_HLD);
      const unsigned src = (unsigned)__cvta_generic_to_shared(stage);
#pragma unroll
      for (int c = 0; c < MS_CL; ++c) hf_bulk_push(hf_mapa(dst, c), src, MS_ROWBYTES, hf_mapa(barAG, c));
    }
  };

  // ---- norms / sums of the private columns over this CTA's rows (sparse_nmf.m:157-160,192): partials to the owners ----
  const int a0 = lane, a1 = 32 + lane;
  const bool has1 = a1 < MS_RA;
  // a stream's private columns are handled by warp == stream index
  const bool own_stream = warp < S && slot_s[warp < S ? warp : 0] >= 0;
  const double* wp_mine = Wp + (size_t)(warp < S ? warp : 0) * MS_RA * MS_ROWS;
  // pair p of atom a sits at byte (base(a) ^ ((a & 7) << 4)) ^ (p << 4): the column base is 512-byte aligned
  const unsigned wx0 = ((unsigned)__cvta_generic_to_shared(wp_mine + (size_t)a0 * MS_ROWS)) ^ ((unsigned)(a0 & 7) << 4);
  const unsigned wx1 = ((unsigned)__cvta_generic_to_shared(wp_mine + (size_t)(has1 ? a1 : 0) * MS_ROWS)) ^ ((unsigned)((has1 ? a1 : 0) & 7) << 4);
  if (own_stream) {
    double s1a = 0.0, s2a = 0.0, s1b = 0.0, s2b = 0.0;
#pragma unroll 8
    for (int p = 0; p < 32; ++p) {
      double2 w0, w1 = make_double2(0.0, 0.0);
      asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(w0.x), "=d"(w0.y) : "r"(wx0 ^ ((unsigned)p << 4)));
      if (has1) asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(w1.x), "=d"(w1.y) : "r"(wx1 ^ ((unsigned)p << 4)));
      s1a += w0.x + w0.y;
      s2a = fma(w0.x, w0.x, fma(w0.y, w0.y, s2a));
      s1b += w1.x + w1.y;
      s2b = fma(w1.x, w1.x, fma(w1.y, w1.y, s2b));
    }
    if (tail_rank) {
      const double t0 = WnP[warp * MS_RA + a0], t1 = has1 ? WnP[warp * MS_RA + a1] : 0.0;
      s1a += t0;
      s2a = fma(t0, t0, s2a);
      s1b += t1;
      s2b = fma(t1, t1, s2b);
    }
    double* row = gpart + (size_t)warp * MS_HLD;
    row[a0] = s1a;
    row[MS_PRIV0 + a0] = s2a;
    if (has1) {
      row[a1] = s1b;
      row[MS_PRIV0 + a1] = s2b;
    }
  }
  push_partials();

  // ---- owner state: thread a < 200 <-> atom a of the row layout (shared atoms first) ----
  // (threads 200..203 fill the two pad atoms, the done flag and the spare slot of the outgoing row)
  const bool is_atom = tid < MS_RS + MS_RA;
  const int ridx = tid < MS_RS ? tid : tid + (MS_PRIV0 - MS_RS);
  // position of the atom in the reference's H: [x | adapted d | fixed d]
  const int orig = tid < MS_RX ? tid : (tid < MS_RS ? tid + MS_RA : tid - MS_RF);
  double h = 0.0, inv_wn = 0.0, dph = 0.0;
  int it = 0;
  bool done = my_slot < 0;
  double last_cost = INFINITY, cost = 0.0;
  auto stage_row = [&](double x) {
    if (is_atom) stage[ridx] = x;
    else if (tid < MS_ROWLEN) {
      const int idx = tid == 200 ? MS_RS : (tid == 201 ? MS_RS + 1 : tid);
      stage[idx] = (idx == MS_FLAG && done) ? 1.0 : 0.0;
    }
  };
  if (rank < S) {
    hf_mbar_wait_bounded(barRS, parRS);
    parRS ^= 1u;
    if (is_atom && my_slot >= 0) {
      double wn, s1;
      if (tid < MS_RS) {
        wn = colstat[tid];
        s1 = colstat[MS_KS + tid];
      } else {
        double s2 = 0.0;
        s1 = 0.0;
#pragma unroll
        for (int c = 0; c < MS_CL; ++c) {
          s2 += recv[(size_t)c * MS_ROWLEN + ridx];
          s1 += recv[(size_t)c * MS_ROWLEN + (tid - MS_RS)];
        }
        wn = sqrt(s2);
      }
      inv_wn = 1.0 / wn;
      dph = 1.0 / fmax(s1 / wn + sc.sparsity, flr);   // reciprocal of the H-update denominator, sparse_nmf.m:192-193
      h = h_init[orig] * wn;                          // :160
    }
    double hs = warp_sum(h);
    if (lane == 0) hsumw[warp] = hs;
    stage_row(h * inv_wn);
    push_row();
  }

  // ---- multiplicative updates ----
  const double* hb = hS + (size_t)li * MS_HLD + 2 * lj;   // B fragments of the Lambda pass
  const double* rb = rS + (size_t)li * MS_RLD + 2 * lj;   // B fragments of the g pass
  const int frow = 8 * warp + li;

  // tail row (row 512) on the last rank, warp 7: lanes <-> atoms, one stream after the other
  auto tail_lambda = [&](int mode) {
#pragma unroll
    for (int n = 0; n < S; ++n) {
      double s = 0.0;
      const double* hr = hS + (size_t)n * MS_HLD;
      for (int a = lane; a < MS_KS; a += 32) {
        const bool is_x = a < MS_RX;
        if (mode == 0 || (mode == 1) == is_x) s = fma(WnS[a], hr[a], s);
      }
      if (mode != 1)
        for (int a = lane; a < MS_RA; a += 32) s = fma(WnP[n * MS_RA + a], hr[MS_PRIV0 + a], s);
      s = warp_sum(s);
      if (lane == 0) lamN[n] = s;
    }
    __syncwarp();
  };

  for (;;) {
    hf_mbar_wait_bounded(barAG, parAG);
    parAG ^= 1u;
    bool all_done = true, mine_done = true;
#pragma unroll
    for (int n = 0; n < S; ++n) {
      const bool dn = hS[(size_t)n * MS_HLD + MS_FLAG] != 0.0;
      all_done &= dn;
      if (n == warp) mine_done = dn;
    }
    if (all_done) break;
    // (A) Lambda = W h: shared columns on the tensor cores, private columns as a mat-vec
    double c0, c1;
    ms_pass_a_shared<0>(Wa, hb, lj, c0, c1);
    if (own_stream && !mine_done) ms_pass_a_private(wp_mine, hS + (size_t)warp * MS_HLD + MS_PRIV0, lane, lam_p + (size_t)warp * MS_LLD);
    if (tail_rank && warp == MS_WARPS - 1) {
      tail_lambda(0);
      double ctn = 0.0;
      if (lane < S) {
        const double lam = fmax(lamN[lane], flr);
        const double r = vN * fast_rcp(lam);
        const bool live = slot_s[lane] >= 0;
        rN[lane] = live ? r : 0.0;
        ctn = live ? fma(vN, fast_log(r, log_tab), lam - vN) : 0.0;
      }
      if (lane < 8) costw[64 + lane] = ctn;
    }
    __syncthreads();
    // (R) ratio and KL terms on the accumulator fragments
    {
      double ct[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int n = 2 * lj + e;
        const bool live = n < S && slot_s[n] >= 0;
        double lam = (e ? c1 : c0) + lam_p[(size_t)n * MS_LLD + frow];
        lam = fmax(lam, flr);
        const double r = v[e] * fast_rcp(lam);
        rS[(size_t)n * MS_RLD + frow] = live ? r : 0.0;
        ct[e] = live ? fma(v[e], fast_log(r, log_tab), lam - v[e]) : 0.0;   // sparse_nmf.m:250
      }
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        double s = ct[e];
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 8);
        s += __shfl_xor_sync(0xffffffffu, s, 16);
        if (li == 0) costw[warp * 8 + 2 * lj + e] = s;
      }
    }
    __syncthreads();
    // (B) g = W' r: partial over this CTA's rows
    {
      double g[3][2];
      if (warp < MS_KT - 16) ms_pass_b_shared<3>(Wb, rb, g); else ms_pass_b_shared<2>(Wb, rb, g);
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        if (warp + 8 * q >= MS_KT) break;
        const int a = 8 * (warp + 8 * q) + li;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          double x = g[q][e];
          if (tail_rank) x = fma(WnS[a], rN[2 * lj + e], x);
          gpart[(size_t)(2 * lj + e) * MS_HLD + a] = x;
        }
      }
    }
    if (own_stream && !mine_done) {
      double x0 = 0.0, x1 = 0.0, y0 = 0.0, y1 = 0.0;
      const double2* rp = reinterpret_cast<const double2*>(rS + (size_t)warp * MS_RLD);
#pragma unroll 8
      for (int p = 0; p < 32; ++p) {
        const double2 rr = rp[p];
        double2 w0, w1 = make_double2(0.0, 0.0);
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(w0.x), "=d"(w0.y) : "r"(wx0 ^ ((unsigned)p << 4)));
        if (has1) asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(w1.x), "=d"(w1.y) : "r"(wx1 ^ ((unsigned)p << 4)));
        x0 = fma(w0.x, rr.x, x0);
        x1 = fma(w0.y, rr.y, x1);
        y0 = fma(w1.x, rr.x, y0);
        y1 = fma(w1.y, rr.y, y1);
      }
      double ga = x0 + x1, gb = y0 + y1;
      if (tail_rank) {
        ga = fma(WnP[warp * MS_RA + a0], rN[warp], ga);
        if (has1) gb = fma(WnP[warp * MS_RA + a1], rN[warp], gb);
      }
      double* row = gpart + (size_t)warp * MS_HLD + MS_PRIV0;
      row[a0] = ga;
      if (has1) row[a1] = gb;
    }
    if (warp == MS_WARPS - 1 && lane < S) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < 9; ++w) s += costw[w * 8 + lane];
      gpart[(size_t)lane * MS_HLD + MS_FLAG] = s;
    }
    push_partials();
    // (C) owner: combine the 8 partials in rank order, stop rule, h update, send h ./ wn back
    if (rank < S) {
      hf_mbar_wait_bounded(barRS, parRS);
      parRS ^= 1u;
      if (!done) {
        double gk = 0.0;
        if (is_atom) {
#pragma unroll
          for (int c = 0; c < MS_CL; ++c) gk += recv[(size_t)c * MS_ROWLEN + ridx];
        }
        bool stop = false;
        if (sc.cost_check && it >= 1) {
          double div = 0.0, hs = 0.0;
#pragma unroll
          for (int c = 0; c < MS_CL; ++c) div += recv[(size_t)c * MS_ROWLEN + MS_FLAG];
#pragma unroll
          for (int w = 0; w < MS_WARPS; ++w) hs += hsumw[(it & 1) * 8 + w];
          cost = div + sc.sparsity * hs;                                   // sparse_nmf.m:261
          if (it > 1 && sc.conv_eps > 0.0) {
            const double e = fabs(cost - last_cost) / last_cost;           // :274
            if (e < sc.conv_eps) stop = true;
          }
          last_cost = cost;
        }
        if (it >= sc.max_iter) stop = true;
        if (stop) {
          done = true;
        } else {
          h = h * (gk * inv_wn) * dph;                                     // :195
          ++it;
          const double hs = warp_sum(h);
          if (lane == 0) hsumw[(it & 1) * 8 + warp] = hs;
        }
      }
      stage_row(h * inv_wn);
      push_row();
    }
  }

  // ---- outputs of the owned stream; activations for the un-normalised basis go back out for the reconstructions ----
  cluster.sync();   // every copy of the last round has landed: the staging rows may be rewritten
  if (tid == 0) hf_mbar_expect_tx(barAG, (unsigned)S * MS_ROWBYTES);
  if (rank < S) {
    if (my_slot >= 0) {
      if (is_atom) st.A[(size_t)my_slot * R + orig] = h;
      if (tid == 0) {
        st.h_iters[my_slot] = it;
        st.h_cost[my_slot] = cost;
      }
    }
    stage_row(h);
    push_row();
  }
  hf_mbar_wait_bounded(barAG, parAG);
  parAG ^= 1u;
  // X_hat = B_x A_x
  {
    double c0, c1;
    ms_pass_a_shared<1>(Wa, hb, lj, c0, c1);
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int n = 2 * lj + e;
      if (n < S && slot_s[n] >= 0) st.Xhat[(size_t)slot_s[n] * LDF + f0 + frow] = e ? c1 : c0;
    }
    if (tail_rank && warp == MS_WARPS - 1) {
      tail_lambda(1);
      if (lane < S && slot_s[lane] >= 0) st.Xhat[(size_t)slot_s[lane] * LDF + FT] = lamN[lane];
      __syncwarp();
    }
  }
  // D_hat = B_d A_d
  {
    double c0, c1;
    ms_pass_a_shared<2>(Wa, hb, lj, c0, c1);
    if (own_stream) ms_pass_a_private(wp_mine, hS + (size_t)warp * MS_HLD + MS_PRIV0, lane, lam_p + (size_t)warp * MS_LLD);
    if (tail_rank && warp == MS_WARPS - 1) {
      tail_lambda(2);
      if (lane < S && slot_s[lane] >= 0) st.Dhat[(size_t)slot_s[lane] * LDF + FT] = lamN[lane];
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int n = 2 * lj + e;
      if (n < S && slot_s[n] >= 0)
        st.Dhat[(size_t)slot_s[n] * LDF + f0 + frow] = (e ? c1 : c0) + lam_p[(size_t)n * MS_LLD + frow];
    }
  }
  cluster.sync();  // nobody may exit while a peer's bulk copy can still target its shared memory
}

// ---- host side ----
constexpr int MS_S = 7;   // streams per cluster: 7 x 25.6 KB of private columns + 43 KB of exchange buffers per CTA

bool hsolve_ms_supported(snmfnat_ctx* ctx, const OnlineDims& d) {
  const int E = d.F - MS_CL * MS_ROWS;
  if (E < 0 || E > 1) return false;
  if (d.R_x != MS_RX || d.R_d != MS_RF + MS_RA || d.R_a != MS_RA || d.R != MS_RX + MS_RF + MS_RA) return false;
  if (d.LDF < d.F || (d.LDF & 1)) return false;
  return (int)MsLayout<MS_S>::bytes <= ctx->max_smem_optin;
}

int hsolve_ms_streams() { return MS_S; }

void launch_ms_colstat(snmfnat_ctx* ctx, const OnlineDims& d, const double* Bx, const double* Bd_fix, double* colstat) {
  ms_colstat_kernel<<<(MS_KS + 7) / 8, 256, 0, ctx->stream>>>(Bx, Bd_fix, d.F, d.LDF, colstat);
  count_launch(ctx);
  check_launch(ctx, "ms_colstat_kernel");
}

void launch_hsolve_ms(snmfnat_ctx* ctx, const OnlineDims& d, const OnlineScalars& sc, const SlotState& st,
                      const FrameArrays& fr, const double* h_init, int n_active, int g_step) {
  SN_REQUIRE(st.ms_colstat != nullptr, SNMFNAT_EINVAL, "multi-stream H-solve: column statistics of the shared basis are missing");
  const size_t bytes = MsLayout<MS_S>::bytes;
  SN_CUDA(cudaFuncSetAttribute(hsolve_ms_kernel<MS_S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  const int groups = (n_active + MS_S - 1) / MS_S;
  hsolve_ms_kernel<MS_S><<<dim3(MS_CL * groups), dim3(MS_THREADS), bytes, ctx->stream>>>(
      d, sc, st, fr, h_init, g_step, log_table(ctx), st.ms_colstat, st.ms_perm, n_active);
  count_launch(ctx);
  check_launch(ctx, "hsolve_ms_kernel");
}

}  // namespace snmfnat
