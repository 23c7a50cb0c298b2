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
//
// Where the private columns live (template <S, TS>): the first S - TS streams in shared memory as described; the last TS
// (one per TMEM lane quadrant, TS <= 4) in TENSOR MEMORY, lane-private, in BOTH access patterns (row pair x atoms: 200
// columns; atom x row pairs for the lane's two atoms: 256 columns), read with tcgen05.ld.32x32b by the warp that owns the
// quadrant.  Tensor memory is otherwise idle in this kernel, delivers ~57 B/clk per quadrant next to the 128 B/clk of
// shared memory (tools/tmem_probe.cu), and frees enough shared memory for the eighth stream: <8, 4> fills the N dimension
// of the mma and halves the shared-memory traffic of the mat-vecs; <7, 0> is the all-shared-memory variant.
#include <cooperative_groups.h>
#include <cmath>
#include <cstdlib>
#include "online.cuh"
#include "online_dev.cuh"
#include "umma.cuh"

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
constexpr int MS_RLD = 72;              // row stride of rS (== 8 mod 16)
constexpr int MS_LLD = 66;              // row stride of lam_p (== 2 mod 8)
constexpr int MS_XB = MS_RX / 8, MS_XR = MS_RX % 8;   // the x | d boundary inside atom tile XB
static_assert(MS_XR % 2 == 0, "the two atoms of a lane's pair must belong to the same class");

// Phase probe (build with -DSNMFNAT_MS_PROBE): thread 0 of rank 0 of cluster 0 accumulates clock64 deltas per phase.
#ifdef SNMFNAT_MS_PROBE
__device__ unsigned long long g_ms_probe[16];
#define MS_TICK(i)                                                       \
  do {                                                                   \
    if (probe) {                                                         \
      const long long t_ = clock64();                                    \
      atomicAdd(&g_ms_probe[i], (unsigned long long)(t_ - tprev));       \
      tprev = t_;                                                        \
    }                                                                    \
  } while (0)
#else
#define MS_TICK(i) do {} while (0)
#endif

constexpr int MS_TM_A = 0;                 // TMEM columns of pattern A: 4 k .. 4 k + 3 = W[2 lane .. +1][k] (two doubles)
constexpr int MS_TM_B = 4 * MS_RA;         // pattern B: MS_TM_B + 8 p .. + 7 = W[2p .. +1][lane], W[2p .. +1][32 + lane]
static_assert(MS_TM_B + 8 * (MS_ROWS / 2) <= 512, "both patterns of a stream must fit the 512 columns of a quadrant");

template <int S, int TS>
struct MsLayout {
  static constexpr int SS = S - TS;                                            // streams whose private columns are in shared memory
  static constexpr size_t off_Wp = 0;                                          // [SS][RA][64] swizzled private columns
  static constexpr size_t off_hS = off_Wp + (size_t)SS * MS_RA * MS_ROWS;      // [8][HLD]  h./wn per stream (+ done flag)
  static constexpr size_t off_recv = off_hS + 8 * MS_HLD;                      // [8][ROWLEN] partials of the owned stream
  static constexpr size_t off_lamp = off_recv + 8 * MS_ROWLEN;                 // [8][LLD] private part of Lambda
  static constexpr size_t off_rS = off_lamp + 8 * MS_LLD;                      // [8][RLD] ratio v./Lambda
  static constexpr size_t off_costw = off_rS + 8 * MS_RLD;                     // [8][8] cost partials per warp + [2][8] tail row
  static constexpr size_t off_WnS = off_costw + 80;                            // [KS] tail row of the shared columns
  static constexpr size_t off_WnP = off_WnS + MS_KS;                           // [S][RA] tail row of the private columns
  static constexpr size_t off_rN = off_WnP + (size_t)((S * MS_RA + 1) & ~1);   // [2][8] ratio of the tail row (by parity)
  static constexpr size_t off_hsum = off_rN + 16;                              // [2][8]
  static constexpr size_t off_bar = off_hsum + 16;                             // 2 mbarriers
  static constexpr size_t off_slot = off_bar + 2;                              // 8 ints, then the TMEM base address
  static constexpr size_t doubles = off_slot + 6;
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

__device__ __forceinline__ double2 ms_lds2(unsigned addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}

// The private atom pairs (k, k+1) are visited class by class (k & 7 = 0, 2, 4, 6) so that the swizzled column addresses of
// a whole class are one XOR-ed base plus compile-time offsets: pair t of the schedule starts at atom ms_pk(t).
__host__ __device__ constexpr int ms_pk(int t) {
  return t < 7 ? 8 * t : (t < 13 ? 2 + 8 * (t - 7) : (t < 19 ? 4 + 8 * (t - 13) : 6 + 8 * (t - 19)));
}
static_assert(MS_RA == 50, "ms_pk covers the 25 atom pairs of R_a = 50");

// Lambda pass of one warp: the 8 x 8 accumulator tile (rows 8 warp .. +7, all streams) of the shared columns on the tensor
// cores, interleaved with the private mat-vec of stream `warp` (lanes <-> row pairs) so that the FP64 pipe and the
// shared-memory pipe are busy together.  MODE 0: all atoms; 1: speech atoms only (B_x A_x); 2: noise atoms only (B_d A_d).
//   hb  : shared-memory address of hS[li][2 lj]              (B fragments: atoms 8u + 2 lj + {0,1} of stream li)
//   wX  : address of the stream's private columns + 16 lane  (row pair `lane` of atom k at (wX ^ ((k & 7) << 4)) + 512 k)
//   hp  : address of hS[warp][PRIV0]
//   PRIV : 0 = no private work, 1 = private columns in shared memory (wX), 2 = in tensor memory (wX = TMEM address of the
//          lane quadrant; pattern A)
__device__ __forceinline__ double ms_u2d(unsigned lo, unsigned hi) { return __hiloint2double((int)hi, (int)lo); }
template <int MODE, int PRIV>
__device__ __forceinline__ void ms_pass_a(const double (&Wa)[2 * MS_KT], unsigned hb, int lj, unsigned wX, unsigned hp,
                                          double& c0, double& c1, double& l0, double& l1) {
  double p0 = 0.0, p1 = 0.0, q0 = 0.0, q1 = 0.0;
  double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
  constexpr int PSTEPS = MS_RA / 2;
#pragma unroll
  for (int u = 0; u < MS_KT; ++u) {
    const bool skip = (MODE == 1 && (u > MS_XB || (u == MS_XB && MS_XR == 0))) || (MODE == 2 && u < MS_XB);
    if (!skip) {
      double2 hh = ms_lds2(hb + 64u * (unsigned)u);
      if (MODE != 0 && u == MS_XB) {
        const bool is_x = 2 * lj < MS_XR;
        if ((MODE == 1) != is_x) hh = make_double2(0.0, 0.0);
      }
      // two accumulator chains, alternating: consecutive mma never depend on each other (26 clk dependent latency)
      dmma884(p0, p1, Wa[2 * u], hh.x);
      dmma884(q0, q1, Wa[2 * u + 1], hh.y);
    }
    if (PRIV == 1) {
#pragma unroll
      for (int t = (u * PSTEPS) / MS_KT; t < ((u + 1) * PSTEPS) / MS_KT; ++t) {
        const int k = ms_pk(t);
        const double2 hh = ms_lds2(hp + 8u * (unsigned)k);
        const double2 w0 = ms_lds2((wX ^ ((unsigned)(k & 7) << 4)) + 512u * (unsigned)k);
        const double2 w1 = ms_lds2((wX ^ ((unsigned)((k + 1) & 7) << 4)) + 512u * (unsigned)(k + 1));
        a0 = fma(w0.x, hh.x, a0);
        a1 = fma(w0.y, hh.x, a1);
        b0 = fma(w1.x, hh.y, b0);
        b1 = fma(w1.y, hh.y, b1);
      }
    }
    if (PRIV == 2) {
      // tensor memory is lane-private: no swizzle, atoms in natural order; the loads of this step's atom pairs were issued
      // before the mma above would have been ideal, but one wait covers them all and the mma hide most of the latency
#pragma unroll
      for (int t = (u * PSTEPS) / MS_KT; t < ((u + 1) * PSTEPS) / MS_KT; ++t) {
        const int k = 2 * t;
        unsigned w[8];
        umma::tmem_ld8(wX + (unsigned)(MS_TM_A + 4 * k), w);
        const double2 hh = ms_lds2(hp + 8u * (unsigned)k);
        umma::tmem_wait_ld();
        a0 = fma(ms_u2d(w[0], w[1]), hh.x, a0);
        a1 = fma(ms_u2d(w[2], w[3]), hh.x, a1);
        b0 = fma(ms_u2d(w[4], w[5]), hh.y, b0);
        b1 = fma(ms_u2d(w[6], w[7]), hh.y, b1);
      }
    }
  }
  c0 = p0 + q0;
  c1 = p1 + q1;
  l0 = a0 + b0;
  l1 = a1 + b1;
}

// g pass of one warp: NT atom tiles (atoms 8 (warp + 8q) .. +7, all streams) over this CTA's 64 rows on the tensor cores,
// interleaved with the private mat-vec of stream `warp` (lanes <-> atoms lane and 32 + lane, row pairs in the order
// p = u + 8 j so that the 4 pairs of a step share one XOR-ed base).
//   rb : address of rS[li][2 lj]     rp : address of rS[warp][0]
//   wx0 / wx1 : (column base of the lane's atom) ^ ((atom & 7) << 4)
//   ct0 / ct1 : the lane's two KL terms v log(v / Lambda) - v + Lambda of the ratio step (sparse_nmf.m:250); they are evaluated
//   here, off the critical path, and summed over the 8 rows of the warp's tile into costw_w[stream]
template <int NT, int PRIV>
__device__ __forceinline__ void ms_pass_b(const double (&Wb)[3][16], unsigned rb, unsigned rp, unsigned wx0, unsigned wx1,
                                          double (&g)[3][2], double& ga, double& gb, const double (&v)[2], const double (&lam)[2],
                                          const double (&rat)[2], const double2* __restrict__ log_tab, double* __restrict__ costw_w,
                                          int li, int lj) {
  double x0 = 0.0, x1 = 0.0, y0 = 0.0, y1 = 0.0;
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    // rat == 0 marks a stream slot without a live stream
    double s = rat[e] > 0.0 ? fma(v[e], fast_log(rat[e], log_tab), lam[e] - v[e]) : 0.0;
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    s += __shfl_xor_sync(0xffffffffu, s, 8);
    s += __shfl_xor_sync(0xffffffffu, s, 16);
    if (li == 0) costw_w[2 * lj + e] = s;
  }
#pragma unroll
  for (int q = 0; q < 3; ++q) g[q][0] = g[q][1] = 0.0;
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const double2 rr = ms_lds2(rb + 64u * (unsigned)u);
#pragma unroll
    for (int q = 0; q < NT; ++q) dmma884(g[q][0], g[q][1], Wb[q][2 * u], rr.x);
#pragma unroll
    for (int q = 0; q < NT; ++q) dmma884(g[q][0], g[q][1], Wb[q][2 * u + 1], rr.y);
    if (PRIV == 1) {
      const unsigned b0 = wx0 ^ ((unsigned)u << 4), b1 = wx1 ^ ((unsigned)u << 4);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double2 r2 = ms_lds2(rp + 16u * (unsigned)(u + 8 * j));
        const double2 w0 = ms_lds2(b0 + 128u * (unsigned)j);
        const double2 w1 = ms_lds2(b1 + 128u * (unsigned)j);
        x0 = fma(w0.x, r2.x, x0);
        x1 = fma(w0.y, r2.y, x1);
        y0 = fma(w1.x, r2.x, y0);
        y1 = fma(w1.y, r2.y, y1);
      }
    }
    if (PRIV == 2) {   // wx0 = TMEM address of the lane quadrant; pattern B: 8 columns per row pair (both atoms of the lane)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int p = 4 * u + j;
        unsigned w[8];
        umma::tmem_ld8(wx0 + (unsigned)(MS_TM_B + 8 * p), w);
        const double2 r2 = ms_lds2(rp + 16u * (unsigned)p);
        umma::tmem_wait_ld();
        x0 = fma(ms_u2d(w[0], w[1]), r2.x, x0);
        x1 = fma(ms_u2d(w[2], w[3]), r2.y, x1);
        y0 = fma(ms_u2d(w[4], w[5]), r2.x, y0);
        y1 = fma(ms_u2d(w[6], w[7]), r2.y, y1);
      }
    }
  }
  ga = x0 + x1;
  gb = y0 + y1;
}

template <int S, int TS>
__global__ void __cluster_dims__(MS_CL, 1, 1) __launch_bounds__(MS_THREADS, 1)
hsolve_ms_kernel(OnlineDims d, OnlineScalars sc, SlotState st, FrameArrays fr, const double* __restrict__ h_init, int g_step,
                 const double2* __restrict__ log_tab, const double* __restrict__ colstat, int n_active) {
  using L = MsLayout<S, TS>;
  constexpr int SS = S - TS;
  static_assert(TS == 0 || (TS <= 4 && SS % 4 == 0), "stream n >= SS uses the TMEM quadrant of warp n: (n - SS) == n % 4");
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

#ifdef SNMFNAT_MS_PROBE
  const bool probe = blockIdx.x == 0 && tid == 0;
  long long tprev = clock64();
#endif
  extern __shared__ __align__(1024) double smem[];
  double* Wp = smem + L::off_Wp;
  double* hS = smem + L::off_hS;
  double* recv = smem + L::off_recv;
  double* lam_p = smem + L::off_lamp;
  double* rS = smem + L::off_rS;
  double* costw = smem + L::off_costw;
  double* WnS = smem + L::off_WnS;
  double* WnP = smem + L::off_WnP;
  double* rN = smem + L::off_rN;
  double* hsumw = smem + L::off_hsum;
  int* slot_s = reinterpret_cast<int*>(smem + L::off_slot);
  const unsigned barRS = (unsigned)__cvta_generic_to_shared(smem + L::off_bar);
  const unsigned barAG = barRS + 8u;

  // ---- the streams of this cluster ----
  if (tid < 8) {
    int slot = -1;
    const int p = grp * S + tid;
    if (tid < S && p < n_active) {
      // launch order written by the previous hop's gain kernel (longest solve first, similar lengths together), if any
      const int pg = d.slot0 & 15;
      const bool sorted = st.ms_perm && d.slot0 < 16 && st.ms_perm_step[pg] == g_step;
      const int idx = sorted ? st.ms_perm[(size_t)pg * st.ms_perm_stride + p] : p;
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

  // ---- tensor memory for the private columns of the last TS streams (all 512 columns: one CTA per SM) ----
  unsigned tq = 0;   // TMEM address of this warp's lane quadrant
  if (TS > 0) {
    unsigned* tbase_s = reinterpret_cast<unsigned*>(slot_s + 8);
    if (warp == 0) umma::tmem_alloc(tbase_s, 512);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    tq = *tbase_s + ((unsigned)(32 * (warp & 3)) << 16);
  }

  // ---- private columns: 16-byte cp.async into the swizzled layout (pair p of atom k at slot p ^ (k & 7)) ----
  for (int c = tid; c < SS * MS_RA * 32; c += MS_THREADS) {
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
  // streams SS .. S-1: warp n writes stream n into its TMEM quadrant, in both access patterns
  if (TS > 0 && warp >= SS && warp < S) {
    const int slot = slot_s[warp];
    const double* Bp = slot >= 0 ? st.Bd[st.bd_sel[slot]] + (size_t)slot * d.R_d * LDF + f0 : nullptr;
    // pattern A: lane <-> row pair, columns 4k .. 4k+3 = W[2 lane .. +1][k]
#pragma unroll 5
    for (int k = 0; k < MS_RA; k += 2) {
      double2 w0 = make_double2(0.0, 0.0), w1 = w0;
      if (Bp) {
        w0 = *reinterpret_cast<const double2*>(Bp + (size_t)k * LDF + 2 * lane);
        w1 = *reinterpret_cast<const double2*>(Bp + (size_t)(k + 1) * LDF + 2 * lane);
      }
      const unsigned r[8] = {(unsigned)__double2loint(w0.x), (unsigned)__double2hiint(w0.x), (unsigned)__double2loint(w0.y),
                             (unsigned)__double2hiint(w0.y), (unsigned)__double2loint(w1.x), (unsigned)__double2hiint(w1.x),
                             (unsigned)__double2loint(w1.y), (unsigned)__double2hiint(w1.y)};
      umma::tmem_st8(tq + (unsigned)(MS_TM_A + 4 * k), r);
    }
    // pattern B: lane <-> atoms lane and 32 + lane, columns 8p .. 8p+7 = rows 2p, 2p+1 of both
    const int b0 = lane, b1 = 32 + lane < MS_RA ? 32 + lane : lane;
#pragma unroll 4
    for (int p = 0; p < MS_ROWS / 2; ++p) {
      double2 w0 = make_double2(0.0, 0.0), w1 = w0;
      if (Bp) {
        w0 = *reinterpret_cast<const double2*>(Bp + (size_t)b0 * LDF + 2 * p);
        w1 = *reinterpret_cast<const double2*>(Bp + (size_t)b1 * LDF + 2 * p);
      }
      const unsigned r[8] = {(unsigned)__double2loint(w0.x), (unsigned)__double2hiint(w0.x), (unsigned)__double2loint(w0.y),
                             (unsigned)__double2hiint(w0.y), (unsigned)__double2loint(w1.x), (unsigned)__double2hiint(w1.x),
                             (unsigned)__double2loint(w1.y), (unsigned)__double2hiint(w1.y)};
      umma::tmem_st8(tq + (unsigned)(MS_TM_B + 8 * p), r);
    }
    umma::tmem_wait_st();
  }

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
  // a stream's private columns (and its tail-row element) are handled by warp == stream index
  const bool own_stream = warp < S && slot_s[warp < S ? warp : 0] >= 0;
  double vN = 1.0;        // V of the tail row of stream `warp`
  if (tail_rank && own_stream) vN = fmax(fr.Ym[(size_t)(st.frame_base[slot_s[warp]] + g_step) * LDF + FT], flr);

  // ---- small buffers ----
  for (int i = tid; i < 8 * MS_HLD; i += MS_THREADS) hS[i] = 0.0;
  for (int i = tid; i < 8 * MS_RLD; i += MS_THREADS) rS[i] = 0.0;
  for (int i = tid; i < 8 * MS_LLD; i += MS_THREADS) lam_p[i] = 0.0;
  if (tid < 80) costw[tid] = 0.0;
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
  MS_TICK(12);
  asm volatile("cp.async.wait_all;" ::: "memory");
  MS_TICK(13);
  cluster.sync();   // barriers initialised everywhere, local buffers staged
  MS_TICK(9);

  // ---- exchange: every value is pushed by the lane that produced it (st.async; the receiver's mbarrier counts the bytes) ----
  unsigned parRS = 0, parAG = 0;
  const unsigned recv_mine = (unsigned)__cvta_generic_to_shared(recv + (size_t)rank * MS_ROWLEN);   // my row at an owner
  const unsigned hS_mine = (unsigned)__cvta_generic_to_shared(hS + (size_t)rank * MS_HLD);          // my row at a peer
  // partial of (atom index idx of the row layout, stream n) to the owner n
  auto push_partial = [&](int n, int idx, double x) {
    hf_st_async(hf_mapa(recv_mine + 8u * (unsigned)idx, (unsigned)n), x, hf_mapa(barRS, (unsigned)n));
  };
  constexpr unsigned RS_BYTES = 8u * (MS_KS + MS_RA + 1) * 8u;    // per round at an owner: 8 ranks x (152 + 50 atoms + cost)
  constexpr unsigned AG_BYTES = (unsigned)S * (MS_RS + MS_RA + 1) * 8u;   // per round at every CTA: S owners x (200 atoms + flag)

  // ---- norms / sums of the private columns over this CTA's rows (sparse_nmf.m:157-160,192): partials to the owners ----
  const int a0 = lane, a1 = 32 + lane;
  const bool has1 = a1 < MS_RA;
  const bool in_tmem = TS > 0 && warp >= SS && warp < S;     // this warp's stream has its private columns in tensor memory
  const double* wp_mine = Wp + (size_t)(warp < SS ? warp : 0) * MS_RA * MS_ROWS;
  // pair p of atom a sits at byte (base(a) ^ ((a & 7) << 4)) ^ (p << 4): the column base is 512-byte aligned.  Lanes without a
  // second atom read the first one again (no predicated loads); their result is never pushed.
  const int a1c = has1 ? a1 : a0;
  const unsigned wx0 = ((unsigned)__cvta_generic_to_shared(wp_mine + (size_t)a0 * MS_ROWS)) ^ ((unsigned)(a0 & 7) << 4);
  const unsigned wx1 = ((unsigned)__cvta_generic_to_shared(wp_mine + (size_t)a1c * MS_ROWS)) ^ ((unsigned)(a1c & 7) << 4);
  if (tid == 0) {
    if (rank < S) hf_mbar_expect_tx(barRS, 8u * 2u * MS_RA * 8u);
    hf_mbar_expect_tx(barAG, AG_BYTES);
  }
  if (warp < S) {
    double s1a = 0.0, s2a = 0.0, s1b = 0.0, s2b = 0.0;
    if (own_stream && in_tmem) {
#pragma unroll 4
      for (int p = 0; p < 32; ++p) {
        unsigned w[8];
        umma::tmem_ld8(tq + (unsigned)(MS_TM_B + 8 * p), w);
        umma::tmem_wait_ld();
        const double x0 = ms_u2d(w[0], w[1]), x1 = ms_u2d(w[2], w[3]), y0 = ms_u2d(w[4], w[5]), y1 = ms_u2d(w[6], w[7]);
        s1a += x0 + x1;
        s2a = fma(x0, x0, fma(x1, x1, s2a));
        s1b += y0 + y1;
        s2b = fma(y0, y0, fma(y1, y1, s2b));
      }
    } else if (own_stream) {
#pragma unroll 8
      for (int p = 0; p < 32; ++p) {
        const double2 w0 = ms_lds2(wx0 ^ ((unsigned)p << 4)), w1 = ms_lds2(wx1 ^ ((unsigned)p << 4));
        s1a += w0.x + w0.y;
        s2a = fma(w0.x, w0.x, fma(w0.y, w0.y, s2a));
        s1b += w1.x + w1.y;
        s2b = fma(w1.x, w1.x, fma(w1.y, w1.y, s2b));
      }
    }
    if (own_stream) {
      if (tail_rank) {
        const double t0 = WnP[warp * MS_RA + a0], t1 = WnP[warp * MS_RA + a1c];
        s1a += t0;
        s2a = fma(t0, t0, s2a);
        s1b += t1;
        s2b = fma(t1, t1, s2b);
      }
    }
    push_partial(warp, a0, s1a);
    push_partial(warp, MS_PRIV0 + a0, s2a);
    if (has1) {
      push_partial(warp, a1, s1b);
      push_partial(warp, MS_PRIV0 + a1, s2b);
    }
  }

  // ---- owner state: thread a < 200 <-> atom a of the row layout (shared atoms first); thread 202 carries the done flag.
  //      Only warps 0..6 take part: each of them pushes after its reads of recv, so a peer can never overwrite a value that
  //      is still to be read ----
  const bool is_atom = tid < MS_RS + MS_RA;
  const bool owner_thread = tid < 7 * 32;
  const int ridx = tid < MS_RS ? tid : tid + (MS_PRIV0 - MS_RS);
  // position of the atom in the reference's H: [x | adapted d | fixed d]
  const int orig = tid < MS_RX ? tid : (tid < MS_RS ? tid + MS_RA : tid - MS_RF);
  double h = 0.0, inv_wn = 0.0, dph = 0.0;
  int it = 0;
  bool done = my_slot < 0;
  double last_cost = INFINITY, cost = 0.0;
  // the owner's row (x for the atoms, the done flag) to all 8 CTAs
  auto push_row = [&](double x) {
    if (is_atom || tid == MS_FLAG) {
      const double val = is_atom ? x : (done ? 1.0 : 0.0);
      const unsigned off = 8u * (unsigned)(is_atom ? ridx : MS_FLAG);
#pragma unroll
      for (int c = 0; c < MS_CL; ++c) hf_st_async(hf_mapa(hS_mine + off, (unsigned)c), val, hf_mapa(barAG, (unsigned)c));
    }
  };
  if (rank < S && owner_thread) {
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
    const double hs = warp_sum(h);
    if (lane == 0) hsumw[warp] = hs;
    push_row(h * inv_wn);
  }

  // ---- multiplicative updates ----
  const unsigned hb = (unsigned)__cvta_generic_to_shared(hS + (size_t)li * MS_HLD + 2 * lj);   // B fragments of the Lambda pass
  const unsigned rb = (unsigned)__cvta_generic_to_shared(rS + (size_t)li * MS_RLD + 2 * lj);   // B fragments of the g pass
  const unsigned hp_mine = (unsigned)__cvta_generic_to_shared(hS + (size_t)(warp < S ? warp : 0) * MS_HLD + MS_PRIV0);
  const unsigned rp_mine = (unsigned)__cvta_generic_to_shared(rS + (size_t)(warp < S ? warp : 0) * MS_RLD);
  const unsigned wX = in_tmem ? tq : (unsigned)__cvta_generic_to_shared(wp_mine) + 16u * (unsigned)lane;
  const unsigned wB0 = in_tmem ? tq : wx0;   // first operand of the g pass: TMEM quadrant or swizzled column base
  double* lamp_mine = lam_p + (size_t)(warp < S ? warp : 0) * MS_LLD;
  const int frow = 8 * warp + li;

  // tail row (row 512, last rank): Lambda of stream `warp`, lanes <-> atom pairs.  mode as in ms_pass_a
  auto tail_lambda = [&](int mode) -> double {
    double s = 0.0;
    const double* hr = hS + (size_t)warp * MS_HLD;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int a = 2 * (lane + 32 * j);
      if (a < MS_KS) {
        const bool is_x = a < MS_RX;
        if (mode == 0 || (mode == 1) == is_x) {
          const double2 w = *reinterpret_cast<const double2*>(WnS + a);
          const double2 x = *reinterpret_cast<const double2*>(hr + a);
          s = fma(w.x, x.x, fma(w.y, x.y, s));
        }
      }
    }
    if (mode != 1 && 2 * lane < MS_RA) {
      const double2 w = *reinterpret_cast<const double2*>(WnP + warp * MS_RA + 2 * lane);
      const double2 x = *reinterpret_cast<const double2*>(hr + MS_PRIV0 + 2 * lane);
      s = fma(w.x, x.x, fma(w.y, x.y, s));
    }
    return warp_sum(s);
  };

  MS_TICK(10);
  // The tail-row ratio / cost of iteration t are written in phase A and read in phase B; the writers of iteration t+1 are
  // only released by the exchange (the owners need this CTA's partials first), which a race checker cannot see: the two
  // small arrays are double-buffered by iteration parity so that the ordering is also visible inside the CTA.
  int par = 0;
  for (;;) {
    par ^= 1;
    hf_mbar_wait_bounded(barAG, parAG);
    MS_TICK(0);
    parAG ^= 1u;
    bool all_done = true, mine_done = true;
#pragma unroll
    for (int n = 0; n < S; ++n) {
      const bool dn = hS[(size_t)n * MS_HLD + MS_FLAG] != 0.0;
      all_done &= dn;
      if (n == warp) mine_done = dn;
    }
    if (all_done) break;
    if (tid == 0) {   // this round's exchanges
      if (rank < S) hf_mbar_expect_tx(barRS, RS_BYTES);
      hf_mbar_expect_tx(barAG, AG_BYTES);
    }
    const bool priv = own_stream && !mine_done;
    // (A) Lambda = W h: shared columns on the tensor cores, private columns as a mat-vec
    double c0, c1;
    if (priv) {
      double l0, l1;
      if (in_tmem) ms_pass_a<0, 2>(Wa, hb, lj, wX, hp_mine, c0, c1, l0, l1);
      else ms_pass_a<0, 1>(Wa, hb, lj, wX, hp_mine, c0, c1, l0, l1);
      *reinterpret_cast<double2*>(lamp_mine + 2 * lane) = make_double2(l0, l1);
    } else {
      double l0, l1;
      ms_pass_a<0, 0>(Wa, hb, lj, wX, hp_mine, c0, c1, l0, l1);
    }
    if (tail_rank && warp < S) {
      double ctn = 0.0, rn = 0.0;
      if (priv) {
        const double lam = fmax(tail_lambda(0), flr);
        rn = vN * fast_rcp(lam);
        ctn = fma(vN, fast_log(rn, log_tab), lam - vN);
      }
      if (lane == 0) {
        rN[par * 8 + warp] = rn;
        costw[64 + par * 8 + warp] = ctn;
      }
    }
    MS_TICK(1);
    __syncthreads();
    MS_TICK(2);
    // (R) ratio and KL terms on the accumulator fragments
    double lamv[2], ratv[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int n = 2 * lj + e;
      const bool live = n < S && slot_s[n] >= 0;
      lamv[e] = fmax((e ? c1 : c0) + lam_p[(size_t)n * MS_LLD + frow], flr);
      ratv[e] = live ? v[e] * fast_rcp(lamv[e]) : 0.0;
      rS[(size_t)n * MS_RLD + frow] = ratv[e];
    }
    __syncthreads();
    MS_TICK(3);
    // (B) g = W' r: partial over this CTA's rows, pushed to the owners straight from the registers
    {
      double g[3][2], ga, gb;
      const bool nt3 = warp < MS_KT - 16;
      if (nt3) {
        // (the three-tile warps 0..2 never own a TMEM stream: SS >= 4 whenever TS > 0)
        if (priv) ms_pass_b<3, 1>(Wb, rb, rp_mine, wx0, wx1, g, ga, gb, v, lamv, ratv, log_tab, costw + warp * 8, li, lj);
        else ms_pass_b<3, 0>(Wb, rb, rp_mine, wx0, wx1, g, ga, gb, v, lamv, ratv, log_tab, costw + warp * 8, li, lj);
      } else {
        if (priv && in_tmem) ms_pass_b<2, 2>(Wb, rb, rp_mine, wB0, wx1, g, ga, gb, v, lamv, ratv, log_tab, costw + warp * 8, li, lj);
        else if (priv) ms_pass_b<2, 1>(Wb, rb, rp_mine, wx0, wx1, g, ga, gb, v, lamv, ratv, log_tab, costw + warp * 8, li, lj);
        else ms_pass_b<2, 0>(Wb, rb, rp_mine, wx0, wx1, g, ga, gb, v, lamv, ratv, log_tab, costw + warp * 8, li, lj);
      }
      // the cost partials of all warps are summed by warp 7: the others only arrive at named barrier 1 (non-blocking)
      if (warp != MS_WARPS - 1) asm volatile("bar.arrive 1, 256;" ::: "memory");
      // the lane's two streams are n = 2 lj and 2 lj + 1: remote addresses of its first atom in their owners' rows
      const unsigned la = recv_mine + 8u * (unsigned)(8 * warp + li);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int n = 2 * lj + e;
        if (n < S) {
          const unsigned ra = hf_mapa(la, (unsigned)n), rbar = hf_mapa(barRS, (unsigned)n);
          const double rn = tail_rank ? rN[par * 8 + n] : 0.0;
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            if (q == 2 && !nt3) break;
            double x = g[q][e];
            if (tail_rank) x = fma(WnS[8 * (warp + 8 * q) + li], rn, x);
            hf_st_async(ra + 512u * (unsigned)q, x, rbar);
          }
        }
      }
      if (warp < S) {
        if (tail_rank) {
          ga = fma(WnP[warp * MS_RA + a0], rN[par * 8 + warp], ga);
          gb = fma(WnP[warp * MS_RA + a1c], rN[par * 8 + warp], gb);
        }
        push_partial(warp, MS_PRIV0 + a0, ga);
        if (has1) push_partial(warp, MS_PRIV0 + a1, gb);
      }
      if (warp == MS_WARPS - 1) {
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (lane < S) {
          double s = 0.0;
#pragma unroll
          for (int w = 0; w < 8; ++w) s += costw[w * 8 + lane];
          s += costw[64 + par * 8 + lane];
          push_partial(lane, MS_FLAG, s);
        }
      }
    }
    MS_TICK(4);
    // (C) owner: combine the 8 partials in rank order, stop rule, h update, send h ./ wn back
    if (rank < S && owner_thread) {
      hf_mbar_wait_bounded(barRS, parRS);
      MS_TICK(5);
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
          for (int w = 0; w < 7; ++w) hs += hsumw[(it & 1) * 8 + w];
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
      push_row(h * inv_wn);
      MS_TICK(6);
#ifdef SNMFNAT_MS_PROBE
      if (probe) atomicAdd(&g_ms_probe[8], 1ull);
#endif
    }
  }

  // ---- outputs of the owned stream; activations for the un-normalised basis go back out for the reconstructions ----
  // (inside the loop an owner can only send after every CTA has passed its previous wait; here nothing orders the
  // rounds, so a cluster barrier keeps the bytes of this round out of a peer's previous phase)
  cluster.sync();
  if (tid == 0) hf_mbar_expect_tx(barAG, AG_BYTES);
  if (rank < S && owner_thread) {
    if (my_slot >= 0) {
      if (is_atom) st.A[(size_t)my_slot * R + orig] = h;
      if (tid == 0) {
        st.h_iters[my_slot] = it;
        st.h_cost[my_slot] = cost;
      }
    }
    push_row(h);
  }
  hf_mbar_wait_bounded(barAG, parAG);
  parAG ^= 1u;
  // X_hat = B_x A_x
  {
    double c0, c1, l0, l1;
    ms_pass_a<1, 0>(Wa, hb, lj, wX, hp_mine, c0, c1, l0, l1);
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int n = 2 * lj + e;
      if (n < S && slot_s[n] >= 0) st.Xhat[(size_t)slot_s[n] * LDF + f0 + frow] = e ? c1 : c0;
    }
    if (tail_rank && own_stream) {
      const double x = tail_lambda(1);
      if (lane == 0) st.Xhat[(size_t)slot_s[warp] * LDF + FT] = x;
    }
  }
  // D_hat = B_d A_d
  {
    double c0, c1, l0, l1;
    if (own_stream) {
      if (in_tmem) ms_pass_a<2, 2>(Wa, hb, lj, wX, hp_mine, c0, c1, l0, l1);
      else ms_pass_a<2, 1>(Wa, hb, lj, wX, hp_mine, c0, c1, l0, l1);
      *reinterpret_cast<double2*>(lamp_mine + 2 * lane) = make_double2(l0, l1);
    } else {
      ms_pass_a<2, 0>(Wa, hb, lj, wX, hp_mine, c0, c1, l0, l1);
    }
    if (tail_rank && own_stream) {
      const double x = tail_lambda(2);
      if (lane == 0) st.Dhat[(size_t)slot_s[warp] * LDF + FT] = x;
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int n = 2 * lj + e;
      if (n < S && slot_s[n] >= 0)
        st.Dhat[(size_t)slot_s[n] * LDF + f0 + frow] = (e ? c1 : c0) + lam_p[(size_t)n * MS_LLD + frow];
    }
  }
  cluster.sync();  // nobody may exit while a peer can still push into its shared memory
  if (TS > 0 && warp == 0) umma::tmem_dealloc(tq, 512);   // warp 0's quadrant starts at lane 0: tq is the address tcgen05.alloc returned
  MS_TICK(11);
#ifdef SNMFNAT_MS_PROBE
  if (probe) atomicAdd(&g_ms_probe[15], 1ull);
#endif
}

// ---- host side ----
// Streams per cluster.  <8, 4>: four streams' private columns in shared memory (102 KB), four in tensor memory, the mma N
// dimension full.  <7, 0> (SNMFNAT_HSOLVE=ms7): all seven in shared memory (179 KB + 43 KB of exchange buffers).
constexpr int MS_S = 8, MS_TS = 4;
static bool ms_use_tmem() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SNMFNAT_HSOLVE");
    v = (e && std::string(e) == "ms7") ? 0 : 1;
  }
  return v == 1;
}

bool hsolve_ms_supported(snmfnat_ctx* ctx, const OnlineDims& d) {
  const int E = d.F - MS_CL * MS_ROWS;
  if (E < 0 || E > 1) return false;
  if (d.R_x != MS_RX || d.R_d != MS_RF + MS_RA || d.R_a != MS_RA || d.R != MS_RX + MS_RF + MS_RA) return false;
  if (d.LDF < d.F || (d.LDF & 1)) return false;
  return (int)MsLayout<7, 0>::bytes <= ctx->max_smem_optin;
}

int hsolve_ms_streams() { return ms_use_tmem() ? MS_S : 7; }

void launch_ms_colstat(snmfnat_ctx* ctx, const OnlineDims& d, const double* Bx, const double* Bd_fix, double* colstat) {
  ms_colstat_kernel<<<(MS_KS + 7) / 8, 256, 0, ctx->stream>>>(Bx, Bd_fix, d.F, d.LDF, colstat);
  count_launch(ctx);
  check_launch(ctx, "ms_colstat_kernel");
}

template <int S, int TS>
static void launch_ms_variant(snmfnat_ctx* ctx, const OnlineDims& d, const OnlineScalars& sc, const SlotState& st,
                              const FrameArrays& fr, const double* h_init, int n_active, int g_step) {
  const size_t bytes = MsLayout<S, TS>::bytes;
  auto kern = hsolve_ms_kernel<S, TS>;
  SN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  const int groups = (n_active + S - 1) / S;
  static bool reported = false;
  if (!reported && getenv("SNMFNAT_DEBUG")) {
    reported = true;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(MS_CL * groups);
    cfg.blockDim = dim3(MS_THREADS);
    cfg.dynamicSmemBytes = bytes;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = MS_CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nc = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&nc, kern, &cfg);
    fprintf(stderr, "snmfnat: hsolve_ms_kernel<%d,%d>: %zu bytes of shared memory, max active clusters = %d (%s)\n", S, TS, bytes, nc,
            cudaGetErrorString(e));
  }
  kern<<<dim3(MS_CL * groups), dim3(MS_THREADS), bytes, ctx->stream>>>(d, sc, st, fr, h_init, g_step, log_table(ctx), st.ms_colstat,
                                                                        n_active);
}

void launch_hsolve_ms(snmfnat_ctx* ctx, const OnlineDims& d, const OnlineScalars& sc, const SlotState& st,
                      const FrameArrays& fr, const double* h_init, int n_active, int g_step) {
  SN_REQUIRE(st.ms_colstat != nullptr, SNMFNAT_EINVAL, "multi-stream H-solve: column statistics of the shared basis are missing");
  if (ms_use_tmem()) launch_ms_variant<MS_S, MS_TS>(ctx, d, sc, st, fr, h_init, n_active, g_step);
  else launch_ms_variant<7, 0>(ctx, d, sc, st, fr, h_init, n_active, g_step);
  count_launch(ctx);
  check_launch(ctx, "hsolve_ms_kernel");
#ifdef SNMFNAT_MS_PROBE
  static int nl = 0;
  if (++nl == 100) {
    SN_CUDA(cudaStreamSynchronize(ctx->stream));
    unsigned long long pr[16];
    SN_CUDA(cudaMemcpyFromSymbol(pr, g_ms_probe, sizeof(pr)));
    const double it = (double)(pr[8] ? pr[8] : 1);
    fprintf(stderr, "snmfnat ms probe (clk per iteration over %.0f iterations): waitAG %.0f | passA %.0f | bar1 %.0f | ratio+bar2 %.0f | "
            "passB+push %.0f | waitRS %.0f | owner+pushrow %.0f\n", it, pr[0] / it, pr[1] / it, pr[2] / it, pr[3] / it, pr[4] / it,
            pr[5] / it, pr[6] / it);
    const double nk = (double)(pr[15] ? pr[15] : 1);
    fprintf(stderr, "snmfnat ms probe (clk per kernel, cluster 0, %.0f kernels): issue staging %.0f | cp.async wait %.0f | cluster.sync %.0f | "
            "init exchange %.0f | epilogue %.0f | iterations per kernel %.1f\n", nk, pr[12] / nk, pr[13] / nk, pr[9] / nk, pr[10] / nk,
            pr[11] / nk, it / nk);
  }
#endif
}

}  // namespace snmfnat
