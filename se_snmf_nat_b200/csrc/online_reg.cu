// Third-generation H-solve for the shipped geometry: the basis lives in REGISTERS.
//
// hsolve_reg_kernel<CT> : one 8-CTA cluster per stream, 8 warps per CTA.  CTA `rank` owns rows [64 rank, 64 rank + 64)
//   of W = [B_x B_d]; warp w owns 8 of them for ALL atoms; lane (rg, cg) holds the 4 x CT tile
//   rows 8w + 4rg + {0..3}, atoms cg CT + {0..CT-1} in registers for every MU iteration of the frame (W never touches
//   shared memory, which is what bounded the second-generation kernel: two reads of the 205 KB slice per iteration).
//     lambda = W h        : 4 CT FMAs per lane, then a butterfly over the 16 cg lanes (5 shuffles) - no barrier
//     r = v ./ lambda     : in the lanes that end up holding the row sums, handed back to the row group by 4 shuffles
//     g = W' r  (partial) : 4 CT FMAs per lane from the same registers, 16 partials per atom through shared memory
//   The per-CTA partial of g (R values + the cost term) is PUSHED to the 8 CTAs of the cluster with st.async
//   (distributed shared memory, completion counted in bytes on the receiver's mbarrier): one block barrier, one
//   mbarrier wait and no cluster barrier / fence per iteration.  Every CTA adds the 8 partials in rank order, so
//   all of them hold bit-identical h, cost and stop decision.
//   Rows >= 512 (the Nyquist bin of F = 513) sit in a small shared-memory side array on the last rank.
//
// Reference: src/sparse_nmf.m:157-208,247-286 (H-update, KL) as called from src/bnmf_sep_event_RT_IS16.m:124-154,
// reconstructions of :158-202.
#include <cooperative_groups.h>
#include <cmath>
#include "online.cuh"

namespace cg = cooperative_groups;

namespace snmfnat {

constexpr int HR_CL = 8;         // CTAs per cluster
constexpr int HR_ROWS = 64;      // rows per CTA
constexpr int HR_THREADS = 256;
constexpr int HR_WARPS = 8;
constexpr int HR_CG = 16;        // column groups (lanes of a row group)
constexpr int HR_NP = 16;        // partials per atom inside a CTA (8 warps x 2 row groups)
constexpr int HR_EMAX = 8;       // rows beyond 512 handled by the side array

template <int CT>
struct HrLayout {
  static constexpr int KP = HR_CG * CT;          // padded atom count; position p = j*16 + cg <-> atom cg*CT + j
  static constexpr int HS = (CT + 1) & ~1;       // per-column-group stride of the h staging (16-byte aligned groups)
  static constexpr int XN = KP + 2;              // exchange row: KP partials + cost partial (+ pad)
  static constexpr size_t off_gp = 0;                                         // [16][KP]
  static constexpr size_t off_recv = off_gp + (size_t)HR_NP * KP;             // [2][8][XN]
  static constexpr size_t off_h = off_recv + (size_t)2 * HR_CL * XN;          // [16][HS]
  static constexpr size_t off_wt = off_h + (size_t)HR_CG * HS;                // [EMAX][KP] tail rows (last rank)
  static constexpr size_t off_misc = off_wt + (size_t)HR_EMAX * KP;           // 64 doubles
  static constexpr size_t off_bar = off_misc + 64;                            // 2 mbarriers
  static constexpr size_t bytes = (off_bar + 2) * sizeof(double);
};

// ---- PTX helpers: mbarrier + st.async over distributed shared memory
__device__ __forceinline__ unsigned hr_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned hr_mapa(unsigned addr, unsigned rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void hr_st_async(unsigned raddr, double v, unsigned rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
               :: "r"(raddr), "l"(__double_as_longlong(v)), "r"(rbar) : "memory");
}
__device__ __forceinline__ void hr_mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void hr_mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void hr_mbar_wait(unsigned bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void hr_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ double hr_rcp(double x) {  // x > 0, normal: <= 1 ulp after two Newton steps
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  return y;
}

template <int CT>
__global__ void __cluster_dims__(HR_CL, 1, 1) __launch_bounds__(HR_THREADS, 1)
hsolve_reg_kernel(OnlineDims d, OnlineScalars sc, SlotState st, FrameArrays fr, const double* __restrict__ h_init,
                  int g_step, int slot0) {
  using L = HrLayout<CT>;
  constexpr int KP = L::KP, HS = L::HS, XN = L::XN;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int slot = d.slot0 + (int)(blockIdx.x / HR_CL) * d.slot_stride + slot0;
  const int l = g_step + 1 - st.l_offset[slot];
  if (l < 1 || l > st.n_hops[slot]) return;  // uniform over the cluster

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rg = lane >> 4, cgi = lane & 15;
  const int F = d.F, R = d.R, R1 = d.R_x, LDF = d.LDF;
  const int E = F - HR_CL * HR_ROWS;
  const bool tail_rank = (rank == HR_CL - 1) && E > 0;
  const double flr = sc.flr;

  extern __shared__ __align__(16) double smem[];
  double* gp = smem + L::off_gp;        // [16][KP] per-(warp, row group) partials, position-major
  double* recv = smem + L::off_recv;    // [2][8][XN]
  double* h_s = smem + L::off_h;        // [16][HS]: h (or another per-atom vector) grouped by column group
  double* Wt = smem + L::off_wt;        // [E][KP] tail rows, position-indexed
  double* misc = smem + L::off_misc;    // [0..7] cost partial per warp, [8] sum(h), [16..23] tail ratio, [24..31] tail lambda
  const unsigned bar0 = hr_smem_u32(smem + L::off_bar);

  const double* __restrict__ W1 = st.Bx;
  const double* __restrict__ W2 = st.Bd[st.bd_sel[slot]] + (size_t)slot * d.R_d * LDF;
  const long long frame = st.frame_base[slot] + g_step;
  const double* __restrict__ V = fr.Ym + (size_t)frame * LDF;

  if (tid == 0) {
    hr_mbar_init(bar0, 1);
    hr_mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  // ---- this lane's 4 x CT tile of W, straight from global memory into registers ----
  const int row0 = rank * HR_ROWS + warp * 8 + rg * 4;
  double w[4][CT];
#pragma unroll
  for (int j = 0; j < CT; ++j) {
    const int k = cgi * CT + j;
    if (k < R) {
      const double* src = (k < R1 ? W1 + (size_t)k * LDF : W2 + (size_t)(k - R1) * LDF) + row0;
      const double2 a = *reinterpret_cast<const double2*>(src);
      const double2 b = *reinterpret_cast<const double2*>(src + 2);
      w[0][j] = a.x; w[1][j] = a.y; w[2][j] = b.x; w[3][j] = b.y;
    } else {
      w[0][j] = w[1][j] = w[2][j] = w[3][j] = 0.0;
    }
  }
  // position owned by this thread in the per-atom stages (tid < KP): p = tid = jp*16 + cp  <->  atom cp*CT + jp
  const int jp = tid >> 4, cp = tid & 15;
  const int kk = cp * CT + jp;
  const bool owner = tid < KP;
  const bool kvalid = owner && kk < R;
  if (tail_rank && owner) {
    for (int e = 0; e < E; ++e) {
      double x = 0.0;
      if (kvalid) x = (kk < R1 ? W1 + (size_t)kk * LDF : W2 + (size_t)(kk - R1) * LDF)[HR_CL * HR_ROWS + e];
      Wt[e * KP + tid] = x;
    }
  }
  // the row whose sum this lane holds after the butterfly: i* = 2*bit3(cg) + bit2(cg)
  const bool b3 = (cgi & 8) != 0, b2 = (cgi & 4) != 0;
  const int istar = (b3 ? 2 : 0) + (b2 ? 1 : 0);
  const double v_own = fmax(V[row0 + istar], flr);                      // sparse_nmf.m:169
  double v_tail = 0.0;
  if (tail_rank && warp < E && lane == 0) v_tail = fmax(V[HR_CL * HR_ROWS + warp], flr);
  hr_cluster_sync();  // mbarriers of every CTA are initialised before anybody pushes

  unsigned rnd = 0;
  // Sum over the rows of the whole cluster of a per-lane, per-atom quantity: partials through shared memory, the
  // CTA's total (+ tail rows; slot KP carries extra_part(), e.g. the cost partial) pushed to all 8 CTAs, totals added
  // in rank order.  Returns the total of position `tid` (threads < KP); *extra receives the total of slot KP.
  auto allreduce_cols = [&](const double (&val)[CT], auto&& tail_term, auto&& extra_part, double* extra) -> double {
    double* gpr = gp + (size_t)(warp * 2 + rg) * KP + cgi;
#pragma unroll
    for (int j = 0; j < CT; ++j) gpr[j * 16] = val[j];
    __syncthreads();
    const unsigned buf = rnd & 1u, parity = (rnd >> 1) & 1u;
    const unsigned bar = bar0 + 8u * buf;
    double* rb = recv + (size_t)buf * HR_CL * XN;
    if (tid == 0) hr_mbar_expect_tx(bar, (unsigned)(HR_CL * (KP + 1) * sizeof(double)));
    if (tid <= KP) {
      double s = 0.0;
      if (owner) {
#pragma unroll
        for (int p = 0; p < HR_NP; ++p) s += gp[(size_t)p * KP + tid];
        if (tail_rank) s += tail_term(tid);
      } else {
        s = extra_part();
      }
      const unsigned la = hr_smem_u32(rb + (size_t)rank * XN + tid);
#pragma unroll
      for (int c = 0; c < HR_CL; ++c) hr_st_async(hr_mapa(la, c), s, hr_mapa(bar, c));
    }
    hr_mbar_wait(bar, parity);
    double tot = 0.0;
    if (owner) {
#pragma unroll
      for (int c = 0; c < HR_CL; ++c) tot += rb[(size_t)c * XN + tid];
    }
    if (extra) {
      double s = 0.0;
#pragma unroll
      for (int c = 0; c < HR_CL; ++c) s += rb[(size_t)c * XN + KP];
      *extra = s;
    }
    ++rnd;
    return tot;
  };
  auto no_extra = [&]() { return 0.0; };
  // the lane's CT values of a per-atom vector staged in h_s (written by the owner threads, after a barrier)
  auto load_cols = [&](double (&x)[CT]) {
    const double* hp = h_s + cgi * HS;
#pragma unroll
    for (int j = 0; j + 1 < CT; j += 2) {
      const double2 t = *reinterpret_cast<const double2*>(hp + j);
      x[j] = t.x; x[j + 1] = t.y;
    }
    if (CT & 1) x[CT - 1] = hp[CT - 1];
  };
  // row sums of W x over the 16 column groups; the lane ends up with the total of row i* of its row group
  auto row_dot = [&](const double (&x)[CT]) -> double {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
    for (int j = 0; j < CT; ++j) {
      a0 = fma(w[0][j], x[j], a0);
      a1 = fma(w[1][j], x[j], a1);
      a2 = fma(w[2][j], x[j], a2);
      a3 = fma(w[3][j], x[j], a3);
    }
    double k0 = b3 ? a2 : a0, s0 = b3 ? a0 : a2;
    double k1 = b3 ? a3 : a1, s1 = b3 ? a1 : a3;
    k0 += __shfl_xor_sync(0xffffffffu, s0, 8);
    k1 += __shfl_xor_sync(0xffffffffu, s1, 8);
    double k = b2 ? k1 : k0, s = b2 ? k0 : k1;
    k += __shfl_xor_sync(0xffffffffu, s, 4);
    k += __shfl_xor_sync(0xffffffffu, k, 2);
    k += __shfl_xor_sync(0xffffffffu, k, 1);
    return k;
  };
  // tail row `e` (last rank): dot product of the row with the per-atom vector staged in h_s, on warp e
  auto tail_dot = [&](int e) -> double {
    double s = 0.0;
    for (int p = lane; p < KP; p += 32) s = fma(Wt[e * KP + p], h_s[(p & 15) * HS + (p >> 4)], s);
    return warp_sum(s);
  };

  // ---- column norms (sparse_nmf.m:158), h scaling (:160), H-update denominators (:192-193) ----
  double tmp[CT];
#pragma unroll
  for (int j = 0; j < CT; ++j)
    tmp[j] = fma(w[0][j], w[0][j], fma(w[1][j], w[1][j], fma(w[2][j], w[2][j], w[3][j] * w[3][j])));
  const double ss = allreduce_cols(tmp, [&](int p) {
    double s = 0.0;
    for (int e = 0; e < E; ++e) s = fma(Wt[e * KP + p], Wt[e * KP + p], s);
    return s; }, no_extra, nullptr);
#pragma unroll
  for (int j = 0; j < CT; ++j) tmp[j] = (w[0][j] + w[1][j]) + (w[2][j] + w[3][j]);
  const double s1 = allreduce_cols(tmp, [&](int p) {
    double s = 0.0;
    for (int e = 0; e < E; ++e) s += Wt[e * KP + p];
    return s; }, no_extra, nullptr);
  double hk = 0.0, dphk = 0.0, wnk = 1.0;
  if (owner) {
    double inv = 0.0;
    if (kvalid) {
      wnk = sqrt(ss);
      inv = 1.0 / wnk;
      dphk = 1.0 / fmax(s1 * inv + sc.sparsity, flr);   // reciprocal of the H-update denominator
      hk = h_init[kk] * wnk;
    }
    h_s[cp * HS + jp] = inv;
    if (tail_rank)
      for (int e = 0; e < E; ++e) Wt[e * KP + tid] *= inv;
  }
  __syncthreads();
  load_cols(tmp);
#pragma unroll
  for (int j = 0; j < CT; ++j) {
    w[0][j] *= tmp[j]; w[1][j] *= tmp[j]; w[2][j] *= tmp[j]; w[3][j] *= tmp[j];   // :159
  }
  __syncthreads();
  if (owner) h_s[cp * HS + jp] = hk;
  __syncthreads();

  // ---- multiplicative updates ----
  int it = 0;
  double last_cost = INFINITY, cost = 0.0;
  const int src_base = lane & 0x13;
  for (;;) {
    const bool want_cost = sc.cost_check && it >= 1;
    {
      double hx[CT];
      load_cols(hx);
      const double lam = fmax(row_dot(hx), flr);                                       // :207
      const double r_own = v_own * hr_rcp(lam);
      double cpart = 0.0;
      if (want_cost && (lane & 3) == 0) cpart = v_own * log(r_own) - v_own + lam;      // :250
      if (warp == HR_WARPS - 1) {  // sum(h) for the sparsity term of the cost, same order on every CTA
        double s = 0.0;
        for (int p = lane; p < KP; p += 32) s += h_s[(p & 15) * HS + (p >> 4)];
        s = warp_sum(s);
        if (lane == 0) misc[8] = s;
      }
      if (tail_rank && warp < E) {
        const double s = tail_dot(warp);
        if (lane == 0) {
          const double lt = fmax(s, flr);
          const double rt = v_tail * hr_rcp(lt);
          misc[16 + warp] = rt;
          if (want_cost) cpart += v_tail * log(rt) - v_tail + lt;
        }
      }
      if (want_cost) {
        cpart = warp_sum(cpart);
        if (lane == 0) misc[warp] = cpart;
      }
      // ratio of the 4 rows of this row group, then the partial of g = W' r over them
      double r4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        r4[i] = __shfl_sync(0xffffffffu, r_own, src_base | ((i >> 1) << 3) | ((i & 1) << 2));
#pragma unroll
      for (int j = 0; j < CT; ++j)
        tmp[j] = fma(w[0][j], r4[0], fma(w[1][j], r4[1], fma(w[2][j], r4[2], w[3][j] * r4[3])));
    }
    double div = 0.0;
    const double gk = allreduce_cols(tmp, [&](int p) {
      double s = 0.0;
      for (int e = 0; e < E; ++e) s = fma(Wt[e * KP + p], misc[16 + e], s);
      return s; }, [&]() {
      double s = 0.0;
      if (want_cost)
        for (int q = 0; q < HR_WARPS; ++q) s += misc[q];
      return s; }, &div);
    bool stop = false;
    if (want_cost) {
      cost = div + sc.sparsity * misc[8];                                  // :261
      if (it > 1 && sc.conv_eps > 0.0) {
        const double e = fabs(cost - last_cost) / last_cost;               // :274
        if (e < sc.conv_eps) stop = true;
      }
      last_cost = cost;
    }
    if (it >= sc.max_iter) stop = true;
    if (stop) break;
    if (owner) {
      hk = hk * gk * dphk;                                                 // :195
      h_s[cp * HS + jp] = hk;
    }
    __syncthreads();
    ++it;
  }

  // ---- outputs: activations, iteration count, reconstructions X^ = B_x A_x, D^ = B_d A_d with the un-normalised
  //      bases (bnmf_sep_event_RT_IS16.m:174,197): activations scaled back by wn ----
  if (rank == 0) {
    if (kvalid) st.A[(size_t)slot * R + kk] = hk;
    if (tid == 0) {
      st.h_iters[slot] = it;
      st.h_cost[slot] = cost;
    }
  }
  for (int part = 0; part < 2; ++part) {
    __syncthreads();
    if (owner) {
      const bool in_part = kvalid && ((part == 0) == (kk < R1));
      h_s[cp * HS + jp] = in_part ? hk * wnk : 0.0;
    }
    __syncthreads();
    double hx[CT];
    load_cols(hx);
    const double s = row_dot(hx);
    double* dst = (part == 0 ? st.Xhat : st.Dhat) + (size_t)slot * LDF;
    if ((lane & 3) == 0) dst[row0 + istar] = s;
    if (tail_rank && warp < E) {
      const double t = tail_dot(warp);
      if (lane == 0) dst[HR_CL * HR_ROWS + warp] = t;
    }
  }
  hr_cluster_sync();  // nobody exits while a peer could still address its shared memory
}

bool hsolve_reg_supported(snmfnat_ctx* ctx, const OnlineDims& d) {
  const int E = d.F - HR_CL * HR_ROWS;
  if (E < 0 || E > HR_EMAX) return false;
  if (d.R > HR_CG * 13 || d.R < 1) return false;
  if (d.LDF % 4 != 0) return false;
  return (int)HrLayout<13>::bytes <= ctx->max_smem_optin;
}

template <int CT>
static void launch_reg_ct(snmfnat_ctx* ctx, const OnlineDims& d, const OnlineScalars& sc, const SlotState& st,
                          const FrameArrays& fr, const double* h_init, int slot0, int n_slots, int g_step,
                          cudaStream_t stream) {
  const size_t bytes = HrLayout<CT>::bytes;
  SN_CUDA(cudaFuncSetAttribute(hsolve_reg_kernel<CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  hsolve_reg_kernel<CT><<<dim3(HR_CL * n_slots), dim3(HR_THREADS), bytes, stream>>>(d, sc, st, fr, h_init, g_step, slot0);
}

void launch_hsolve_reg(snmfnat_ctx* ctx, const OnlineDims& d, const OnlineScalars& sc, const SlotState& st,
                       const FrameArrays& fr, const double* h_init, int n_active, int g_step) {
  const int ct = (d.R + HR_CG - 1) / HR_CG;
  if (ct <= 4) launch_reg_ct<4>(ctx, d, sc, st, fr, h_init, 0, n_active, g_step, ctx->stream);
  else if (ct <= 7) launch_reg_ct<7>(ctx, d, sc, st, fr, h_init, 0, n_active, g_step, ctx->stream);
  else if (ct <= 10) launch_reg_ct<10>(ctx, d, sc, st, fr, h_init, 0, n_active, g_step, ctx->stream);
  else launch_reg_ct<13>(ctx, d, sc, st, fr, h_init, 0, n_active, g_step, ctx->stream);
  count_launch(ctx);
  check_launch(ctx, "hsolve_reg_kernel");
}

}  // namespace snmfnat
