// Per-hop kernels of the online SNMF-NAT path (float64, sm_100a).
//
//   hsolve_kernel  : sparse_nmf H-update loop for one frame (src/sparse_nmf.m:157-208,247-286) fused with the
//                    per-class reconstruction of src/bnmf_sep_event_RT_IS16.m:158-202.  One 4-CTA cluster per
//                    stream; the basis [B_x B_d] is row-split over the cluster and stays in shared memory for all
//                    iterations; the R-vector W'(v./lambda) is reduced over distributed shared memory.
//   gain_kernel    : src/blk_sparse.m:3-36 + gain / noise-PSD smoothing / adaptation gate and history
//                    (src/bnmf_sep_event_RT_IS16.m:213-292).  One CTA per stream.
//   wsolve_kernel  : sparse_nmf W-update loop on the 100-frame noise history (src/sparse_nmf.m:212-244 called from
//                    src/bnmf_sep_event_RT_IS16.m:322-338).  One 4-CTA cluster per stream, FP64 tensor-core MMAs
//                    (mma.sync m8n8k4.f64) with W and R*H' tiles held in registers.
#include <cooperative_groups.h>
#include <cmath>
#include <cstdlib>
#include "online.cuh"
#include "blk_sparse.cuh"

namespace cg = cooperative_groups;

namespace snmfnat {

// =====================================================================================================
// H-solve
// =====================================================================================================
constexpr int HS_THREADS = 256;
constexpr int HS_WARPS = HS_THREADS / 32;
constexpr int HS_CL = 4;  // CTAs per cluster

struct HsLayout {
  int RS;  // rows per CTA (padded, same on every rank)
  size_t off_W, off_v, off_r, off_lam, off_h, off_dph, off_wn, off_xch, off_scratch, bytes;
  int xn;  // doubles per exchange buffer
};
__host__ __device__ inline HsLayout hs_layout(int F, int R) {
  HsLayout L;
  L.RS = (F + HS_CL - 1) / HS_CL;
  L.xn = 2 * R + 8;
  size_t o = 0;
  L.off_W = o;       o += (size_t)R * L.RS;
  L.off_v = o;       o += L.RS;
  L.off_r = o;       o += L.RS;
  L.off_lam = o;     o += (size_t)HS_WARPS * L.RS;
  L.off_h = o;       o += R;
  L.off_dph = o;     o += R;
  L.off_wn = o;      o += R;
  L.off_xch = o;     o += 2 * (size_t)L.xn;
  L.off_scratch = o; o += 64;
  L.bytes = o * sizeof(double);
  return L;
}
// what the H-solve needs per CTA: the cluster-resident layout when it fits, else the streaming fallback's vectors
size_t hsolve_smem_bytes(const OnlineDims& d) {
  const size_t resident = hs_layout(d.F, d.R).bytes, streamed = ((size_t)5 * d.R + 2 * d.F + 64) * sizeof(double);
  return resident < streamed ? resident : streamed;
}

// sum over the 32 lanes of p[i] for every i; lane L returns the total of p[L]
__device__ __forceinline__ double transpose_reduce32(double (&p)[32], int lane) {
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool up = (lane & half) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const double send = up ? p[i] : p[i + half];
      const double keep = up ? p[i + half] : p[i];
      p[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
  return p[0];
}

template <int RPL>
__device__ __forceinline__ void hs_matvec(const double* __restrict__ Ws, const double* __restrict__ coef, int k_lo,
                                          int k_hi, int RS, int warp, int lane, double* __restrict__ lam_part) {
  double acc[RPL];
#pragma unroll
  for (int j = 0; j < RPL; ++j) acc[j] = 0.0;
  for (int k = k_lo + warp; k < k_hi; k += HS_WARPS) {
    const double hk = coef[k];
    const double* wk = Ws + (size_t)k * RS + lane;
#pragma unroll
    for (int j = 0; j < RPL; ++j)
      if (lane + 32 * j < RS) acc[j] = fma(wk[32 * j], hk, acc[j]);
  }
#pragma unroll
  for (int j = 0; j < RPL; ++j)
    if (lane + 32 * j < RS) lam_part[warp * RS + lane + 32 * j] = acc[j];
}

template <int RPL>
__global__ void __cluster_dims__(HS_CL, 1, 1) __launch_bounds__(HS_THREADS, 1)
hsolve_kernel(OnlineDims d, OnlineScalars sc, SlotState st, FrameArrays fr, const double* __restrict__ h_init,
              int g_step) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int slot = d.slot0 + (int)(blockIdx.x / HS_CL) * d.slot_stride;
  const int l = g_step + 1 - st.l_offset[slot];
  if (l < 1 || l > st.n_hops[slot]) return;  // uniform over the cluster

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int F = d.F, R = d.R, R1 = d.R_x, LDF = d.LDF;
  const HsLayout L = hs_layout(F, R);
  const int RS = L.RS;
  const int base = F / HS_CL, rem = F % HS_CL;
  const int rows = base + (rank < rem ? 1 : 0);
  const int f0 = rank * base + (rank < rem ? rank : rem);

  extern __shared__ __align__(16) double smem[];
  double* Ws = smem + L.off_W;
  double* v_s = smem + L.off_v;
  double* r_s = smem + L.off_r;
  double* lam_part = smem + L.off_lam;
  double* h_s = smem + L.off_h;
  double* dph_s = smem + L.off_dph;
  double* wn_s = smem + L.off_wn;
  double* xch = smem + L.off_xch;
  double* scratch = smem + L.off_scratch;

  const double* __restrict__ W1 = st.Bx;
  const double* __restrict__ W2 = st.Bd[st.bd_sel[slot]] + (size_t)slot * d.R_d * LDF;
  const long long frame = st.frame_base[slot] + g_step;
  const double* __restrict__ V = fr.Ym + (size_t)frame * LDF;
  const double flr = sc.flr;

  // ---- stage the row slice of W = [B_x B_d] (column k contiguous over rows) ----
  for (int k = warp; k < R; k += HS_WARPS) {
    const double* src = (k < R1 ? W1 + (size_t)k * LDF : W2 + (size_t)(k - R1) * LDF) + f0;
    double s1 = 0.0, s2 = 0.0;
    for (int f = lane; f < RS; f += 32) {
      const double x = (f < rows) ? src[f] : 0.0;
      Ws[(size_t)k * RS + f] = x;
      s1 += x;
      s2 = fma(x, x, s2);
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) {
      xch[k] = s2;      // buffer 0: partial sum of squares, partial column sums
      xch[R + k] = s1;
    }
  }
  for (int f = tid; f < RS; f += HS_THREADS) v_s[f] = (f < rows) ? fmax(V[f0 + f], flr) : 0.0;  // sparse_nmf.m:169
  cluster.sync();

  // ---- column norms, h scaling (sparse_nmf.m:157-160), denominators (:192-193) ----
  double hsum = 0.0;
  {
    double hk = 0.0;
    if (tid < R) {
      double s2 = 0.0, s1 = 0.0;
      for (int c = 0; c < HS_CL; ++c) {
        const double* rx = cluster.map_shared_rank(xch, c);
        s2 += rx[tid];
        s1 += rx[R + tid];
      }
      const double wn = sqrt(s2);
      wn_s[tid] = wn;
      dph_s[tid] = fmax(s1 / wn + sc.sparsity, flr);
      hk = h_init[tid] * wn;
      h_s[tid] = hk;
    }
    hsum = block_sum(hk, scratch);
  }
  for (int k = warp; k < R; k += HS_WARPS) {
    const double wn = wn_s[k];
    for (int f = lane; f < RS; f += 32) Ws[(size_t)k * RS + f] = Ws[(size_t)k * RS + f] / wn;
  }
  __syncthreads();

  // ---- multiplicative updates ----
  int it = 0, buf = 1;
  double last_cost = INFINITY, cost = 0.0;
  for (;;) {
    // lambda = max(W*h, flr)                                          sparse_nmf.m:167 / :207
    hs_matvec<RPL>(Ws, h_s, 0, R, RS, warp, lane, lam_part);
    __syncthreads();
    double cterm = 0.0;
    if (tid < RS) {
      double rr = 0.0;
      if (tid < rows) {
        double lam = 0.0;
#pragma unroll
        for (int w = 0; w < HS_WARPS; ++w) lam += lam_part[w * RS + tid];
        lam = fmax(lam, flr);
        const double v = v_s[tid];
        rr = v / lam;
        if (sc.cost_check && it >= 1) cterm = v * log(rr) - v + lam;  // :250
      }
      r_s[tid] = rr;
    }
    const double cost_part = block_sum(cterm, scratch);  // contains the barrier that publishes r_s

    // partial of W'*(v./lambda) over this CTA's rows                  :194
    double rj[RPL];
#pragma unroll
    for (int j = 0; j < RPL; ++j) rj[j] = (lane + 32 * j < RS) ? r_s[lane + 32 * j] : 0.0;
    double* xb = xch + (size_t)buf * L.xn;
    for (int kc = 0; kc < R; kc += 32 * HS_WARPS) {
      double p[32];
#pragma unroll
      for (int kk = 0; kk < 32; ++kk) {
        const int k = kc + kk * HS_WARPS + warp;
        double a = 0.0;
        if (k < R) {
          const double* wk = Ws + (size_t)k * RS + lane;
#pragma unroll
          for (int j = 0; j < RPL; ++j)
            if (lane + 32 * j < RS) a = fma(wk[32 * j], rj[j], a);
        }
        p[kk] = a;
      }
      const double tot = transpose_reduce32(p, lane);
      const int k = kc + lane * HS_WARPS + warp;
      if (k < R) xb[k] = tot;
    }
    if (tid == 0) xb[2 * R] = cost_part;
    cluster.sync();

    // combine the partials of the 4 CTAs in rank order (bit-identical on every CTA)
    double gk = 0.0;
    if (tid < R)
      for (int c = 0; c < HS_CL; ++c) gk += cluster.map_shared_rank(xb, c)[tid];
    bool stop = false;
    if (sc.cost_check && it >= 1) {
      double div = 0.0;
      for (int c = 0; c < HS_CL; ++c) div += cluster.map_shared_rank(xb, c)[2 * R];
      cost = div + sc.sparsity * hsum;                                 // :261
      if (it > 1 && sc.conv_eps > 0.0) {
        const double e = fabs(cost - last_cost) / last_cost;           // :274
        if (e < sc.conv_eps) stop = true;
      }
      last_cost = cost;
    }
    if (it >= sc.max_iter) stop = true;
    if (stop) break;
    double hn = 0.0;
    if (tid < R) {
      hn = h_s[tid] * gk / dph_s[tid];                                 // :195
      h_s[tid] = hn;
    }
    hsum = block_sum(hn, scratch);
    buf ^= 1;
    ++it;
  }

  // ---- outputs: A, per-class reconstruction with the un-normalised basis (bnmf_sep_event_RT_IS16.m:174,197)
  __syncthreads();
  if (tid < R) {
    if (rank == 0) st.A[(size_t)slot * R + tid] = h_s[tid];
    dph_s[tid] = h_s[tid] * wn_s[tid];
  }
  if (rank == 0 && tid == 0) {
    st.h_iters[slot] = it;
    st.h_cost[slot] = cost;
  }
  __syncthreads();
  for (int part = 0; part < 2; ++part) {
    hs_matvec<RPL>(Ws, dph_s, part == 0 ? 0 : R1, part == 0 ? R1 : R, RS, warp, lane, lam_part);
    __syncthreads();
    if (tid < rows) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < HS_WARPS; ++w) s += lam_part[w * RS + tid];
      (part == 0 ? st.Xhat : st.Dhat)[(size_t)slot * LDF + f0 + tid] = s;
    }
    __syncthreads();
  }
  cluster.sync();  // nobody may exit while a peer can still read its exchange buffers
}


// =====================================================================================================
// H-solve for dictionaries that do not fit the shared memory of a 4-CTA cluster (R_x + R_d up to ~4000: the reference's
// exemplar settings use 500 + 500, settings/bak_IS16_results/initial_setting_Exemplar.m:47-48): one CTA per stream, the
// un-normalised basis is STREAMED from L2 / HBM twice per iteration (coalesced column reads), the column scaling of
// sparse_nmf.m:157-160 is folded into the vectors (h / wn on the way in, g / wn on the way out).  Same update, cost and
// stop rule as the resident kernels (sparse_nmf.m:186-283); a fallback, not a throughput path.
// =====================================================================================================
constexpr int HG_THREADS = 576;   // >= F for F = 513; thread <-> row in the Lambda pass
constexpr int HG_WARPS = HG_THREADS / 32;

__global__ void __launch_bounds__(HG_THREADS)
hsolve_stream_kernel(OnlineDims d, OnlineScalars sc, SlotState st, FrameArrays fr, const double* __restrict__ h_init, int g_step) {
  const int slot = d.slot0 + (int)blockIdx.x * d.slot_stride;
  const int l = g_step + 1 - st.l_offset[slot];
  if (l < 1 || l > st.n_hops[slot]) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int F = d.F, R = d.R, R1 = d.R_x, LDF = d.LDF;
  const double flr = sc.flr;
  extern __shared__ __align__(16) double smem[];
  double* h_s = smem;              // [R] activations (normalised-basis convention, sparse_nmf.m:160)
  double* ht_s = h_s + R;          // [R] h ./ wn (what multiplies the raw columns)
  double* iw_s = ht_s + R;         // [R] 1 / wn
  double* dp_s = iw_s + R;         // [R] reciprocal of the H-update denominator (:192-193)
  double* g_s = dp_s + R;          // [R]
  double* v_s = g_s + R;           // [F]
  double* r_s = v_s + F;           // [F]
  double* scratch = r_s + F;       // [64]
  const double* __restrict__ W1 = st.Bx;
  const double* __restrict__ W2 = st.Bd[st.bd_sel[slot]] + (size_t)slot * d.R_d * LDF;
  auto col = [&](int k) { return k < R1 ? W1 + (size_t)k * LDF : W2 + (size_t)(k - R1) * LDF; };
  const double* __restrict__ V = fr.Ym + (size_t)(st.frame_base[slot] + g_step) * LDF;

  for (int k = warp; k < R; k += HG_WARPS) {
    const double* c = col(k);
    double s1 = 0.0, s2 = 0.0;
    for (int f = lane; f < F; f += 32) {
      const double x = c[f];
      s1 += x;
      s2 = fma(x, x, s2);
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) {
      const double wn = sqrt(s2);
      iw_s[k] = 1.0 / wn;
      dp_s[k] = 1.0 / fmax(s1 / wn + sc.sparsity, flr);
      h_s[k] = h_init[k] * wn;
    }
  }
  for (int f = tid; f < F; f += HG_THREADS) v_s[f] = fmax(V[f], flr);          // sparse_nmf.m:169
  __syncthreads();

  // out[f] = sum_{k in [k0, k1)} W[f][k] * x[k] for this thread's row (coalesced over the threads of a warp)
  auto row_dot = [&](const double* x, int k0, int k1, int f) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int k = k0;
    for (; k + 4 <= k1; k += 4) {
      a0 = fma(col(k)[f], x[k], a0);
      a1 = fma(col(k + 1)[f], x[k + 1], a1);
      a2 = fma(col(k + 2)[f], x[k + 2], a2);
      a3 = fma(col(k + 3)[f], x[k + 3], a3);
    }
    for (; k < k1; ++k) a0 = fma(col(k)[f], x[k], a0);
    return (a0 + a1) + (a2 + a3);
  };
  int it = 0;
  double last_cost = INFINITY, cost = 0.0;
  for (;;) {
    for (int k = tid; k < R; k += HG_THREADS) ht_s[k] = h_s[k] * iw_s[k];
    __syncthreads();
    double cterm = 0.0;
    for (int f = tid; f < F; f += HG_THREADS) {
      const double lam = fmax(row_dot(ht_s, 0, R, f), flr);
      const double v = v_s[f];
      r_s[f] = v / lam;
      if (sc.cost_check && it >= 1) cterm += v * log(v / lam) - v + lam;        // :250
    }
    double hpart = 0.0;
    for (int k = tid; k < R; k += HG_THREADS) hpart += h_s[k];
    const double div = block_sum(cterm, scratch);      // also publishes r_s (block barriers inside)
    const double hsum = block_sum(hpart, scratch);
    for (int k = warp; k < R; k += HG_WARPS) {
      const double* c = col(k);
      double s = 0.0;
      for (int f = lane; f < F; f += 32) s = fma(c[f], r_s[f], s);
      s = warp_sum(s);
      if (lane == 0) g_s[k] = s * iw_s[k];
    }
    bool stop = false;
    if (sc.cost_check && it >= 1) {
      cost = div + sc.sparsity * hsum;                                           // :261
      if (it > 1 && sc.conv_eps > 0.0) {
        const double e = fabs(cost - last_cost) / last_cost;                     // :274
        if (e < sc.conv_eps) stop = true;
      }
      last_cost = cost;
    }
    if (it >= sc.max_iter) stop = true;
    if (stop) break;
    __syncthreads();
    for (int k = tid; k < R; k += HG_THREADS) h_s[k] = h_s[k] * g_s[k] * dp_s[k];  // :195
    ++it;
    __syncthreads();
  }
  __syncthreads();
  for (int k = tid; k < R; k += HG_THREADS) st.A[(size_t)slot * R + k] = h_s[k];
  if (tid == 0) {
    st.h_iters[slot] = it;
    st.h_cost[slot] = cost;
  }
  // reconstructions with the un-normalised bases (bnmf_sep_event_RT_IS16.m:174,197)
  for (int f = tid; f < F; f += HG_THREADS) {
    st.Xhat[(size_t)slot * LDF + f] = row_dot(h_s, 0, R1, f);
    st.Dhat[(size_t)slot * LDF + f] = row_dot(h_s, R1, R, f);
  }
}

static size_t hsolve_stream_smem(const OnlineDims& d) { return ((size_t)5 * d.R + 2 * d.F + 64) * sizeof(double); }

bool force_generic() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SNMFNAT_FORCE_GENERIC");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

int hsolve_ms_mode() {
  static int v = -2;
  if (v == -2) {
    const char* e = getenv("SNMFNAT_HSOLVE");
    // "ms" forces the multi-stream kernel for any number of streams, "single" disables it; "ms7" (read in online_ms.cu)
    // selects its all-shared-memory 7-stream variant instead of the 8-stream one with tensor memory
    v = !e ? 0 : (std::string(e) == "ms" ? 1 : (std::string(e) == "single" ? -1 : 0));
  }
  return v;
}

void launch_hsolve(snmfnat_ctx* ctx, const OnlineDims& d, const OnlineScalars& sc, const SlotState& st,
                   const FrameArrays& fr, const double* h_init, int n_active, int g_step) {
  if (n_active <= 0) return;
  if (d.upd1 > d.upd0) {   // p.basis_update_N / _E: the solve also updates part of the dictionary
    launch_hsolve_semi(ctx, d, sc, st, fr, h_init, n_active, g_step);
    return;
  }
  if (!force_generic() && st.ms_colstat && hsolve_ms_supported(ctx, d)) {
    const int mode = hsolve_ms_mode();
    if (mode == 1 || (mode == 0 && n_active >= hsolve_ms_streams())) {
      launch_hsolve_ms(ctx, d, sc, st, fr, h_init, n_active, g_step);
      return;
    }
  }
  if (!force_generic() && hsolve_fast_supported(ctx, d)) {
    launch_hsolve_fast(ctx, d, sc, st, fr, h_init, n_active, g_step);
    return;
  }
  const HsLayout L = hs_layout(d.F, d.R);
  if (d.R > HS_THREADS || (int)L.bytes > ctx->max_smem_optin || (L.RS + 31) / 32 > 5) {
    // dictionary too large for the cluster-resident kernels: stream it (Exemplar / Techwin settings of the reference)
    const size_t sm = hsolve_stream_smem(d);
    SN_REQUIRE((int)sm <= ctx->max_smem_optin, SNMFNAT_EUNSUPPORTED, "H-solve: R_x+R_d = %d needs %zu bytes of shared memory per CTA", d.R, sm);
    if (sm > 48 * 1024) SN_CUDA(cudaFuncSetAttribute(hsolve_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    hsolve_stream_kernel<<<dim3(n_active), dim3(HG_THREADS), sm, ctx->stream>>>(d, sc, st, fr, h_init, g_step);
    count_launch(ctx);
    check_launch(ctx, "hsolve_stream_kernel");
    return;
  }
  SN_REQUIRE(d.R <= HS_THREADS, SNMFNAT_EUNSUPPORTED, "H-solve kernel supports R_x+R_d <= %d (got %d)", HS_THREADS, d.R);
  SN_REQUIRE((int)L.bytes <= ctx->max_smem_optin, SNMFNAT_EUNSUPPORTED,
             "basis slice (%zu bytes) does not fit the %d-byte shared memory of one CTA", L.bytes, ctx->max_smem_optin);
  const int rpl = (L.RS + 31) / 32;
  SN_REQUIRE(rpl <= 5, SNMFNAT_EUNSUPPORTED, "H-solve kernel supports up to 640 frequency rows (got %d)", d.F);
  void (*kern)(OnlineDims, OnlineScalars, SlotState, FrameArrays, const double*, int) = nullptr;
  switch (rpl) {
    case 1: kern = hsolve_kernel<1>; break;
    case 2: kern = hsolve_kernel<2>; break;
    case 3: kern = hsolve_kernel<3>; break;
    case 4: kern = hsolve_kernel<4>; break;
    default: kern = hsolve_kernel<5>; break;
  }
  SN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.bytes));
  kern<<<dim3(HS_CL * n_active), dim3(HS_THREADS), L.bytes, ctx->stream>>>(d, sc, st, fr, h_init, g_step);
  count_launch(ctx);
  check_launch(ctx, "hsolve_kernel");
}

// =====================================================================================================
// gain / block sparsity / adaptation gate
// =====================================================================================================
constexpr int GN_THREADS = 256;

// The last CTA of a gain launch orders the launch's slots for the next hop's multi-stream H-solve: counting sort by this
// hop's iteration count, longest first (consecutive hops are strongly correlated, so streams that share a cluster stop
// at about the same time and the longest solves start first); slots that are inactive at the next hop go last.
__device__ void gain_order_next(const OnlineDims& d, const SlotState& st, int g_step, int n_next) {
  __shared__ int hist[256], base[256];
  __shared__ int last;
  const int tid = threadIdx.x, n = (int)gridDim.x, grp = d.slot0 & 15;
  __syncthreads();   // this CTA's h_iters / state writes are done
  if (tid == 0) {
    __threadfence();
    last = atomicAdd(&st.ms_ticket[grp], 1) == n - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  for (int k = tid; k < 256; k += blockDim.x) hist[k] = 0;
  __syncthreads();
  auto key_of = [&](int i) {
    if (i >= n_next) return 0;
    const int it = st.h_iters[d.slot0 + i * d.slot_stride];
    return 1 + (it < 0 ? 0 : (it > 254 ? 254 : it));
  };
  for (int i = tid; i < n; i += blockDim.x) atomicAdd(&hist[key_of(i)], 1);
  __syncthreads();
  for (int k = tid; k < 256; k += blockDim.x) {
    int b = 0;
    for (int j = k + 1; j < 256; ++j) b += hist[j];
    base[k] = b;
  }
  __syncthreads();
  int* perm = st.ms_perm + (size_t)grp * st.ms_perm_stride;
  for (int i = tid; i < n; i += blockDim.x) perm[atomicAdd(&base[key_of(i)], 1)] = i;
  // the W-solve of THIS hop (launched next on the same stream): gated slots by expected length, the others last
  if (st.ws_perm && st.w_last) {
    __syncthreads();
    for (int k = tid; k < 256; k += blockDim.x) hist[k] = 0;
    __syncthreads();
    auto key_w = [&](int i) {
      const int slot = d.slot0 + i * d.slot_stride;
      if (!st.do_update[slot]) return 0;
      const int kt = (st.n_up[slot] + 7) / 8, passes = st.w_last[slot] > 0 ? st.w_last[slot] + 1 : 6;
      const int k = passes * kt;
      return 1 + (k > 254 ? 254 : k);
    };
    for (int i = tid; i < n; i += blockDim.x) atomicAdd(&hist[key_w(i)], 1);
    __syncthreads();
    for (int k = tid; k < 256; k += blockDim.x) {
      int b = 0;
      for (int j = k + 1; j < 256; ++j) b += hist[j];
      base[k] = b;
    }
    __syncthreads();
    int* pw = st.ws_perm + (size_t)grp * st.ms_perm_stride;
    for (int i = tid; i < n; i += blockDim.x) pw[atomicAdd(&base[key_w(i)], 1)] = i;
    if (tid == 0) st.ws_perm_step[grp] = g_step;
  }
  if (tid == 0) {
    st.ms_ticket[grp] = 0;
    st.ms_perm_step[grp] = g_step + 1;
  }
}

__global__ void __launch_bounds__(GN_THREADS)
gain_kernel(OnlineDims d, OnlineScalars sc, SlotState st, FrameArrays fr, TraceArrays tr, int has_trace, int g_step, int n_next) {
  const int slot = d.slot0 + (int)blockIdx.x * d.slot_stride;
  const int l = g_step + 1 - st.l_offset[slot];
  if (l < 1 || l > st.n_hops[slot]) {
    if (n_next >= 0) gain_order_next(d, st, g_step, n_next);
    return;
  }
  const int tid = threadIdx.x;
  const int F = d.F, LDF = d.LDF, R = d.R, R_x = d.R_x, R_d = d.R_d, R_a = d.R_a, m_a = d.m_a, PL = d.P_len_l;
  const double flr = sc.flr;

  extern __shared__ __align__(16) double smem[];
  double* Q_s = smem;            // [F]
  double* rs1 = Q_s + F;         // [F] row sums of r_blk (also SNR scratch)
  double* rs2 = rs1 + F;         // [F] row sums of squares
  double* G_s = rs2 + F;         // [F]
  double* P_s = G_s + F;         // [F] window values
  double* scratch = P_s + F;     // [64]
  int* up_s = reinterpret_cast<int*>(scratch + 64);  // [R_a]

  const double* A = st.A + (size_t)slot * R;
  const double* Xh = st.Xhat + (size_t)slot * LDF;
  const double* Dh = st.Dhat + (size_t)slot * LDF;
  const long long frame = st.frame_base[slot] + g_step;
  const double* Ym = fr.Ym + (size_t)frame * LDF;
  double* Xt = fr.Xt + (size_t)frame * LDF;

  // activation means (bnmf_sep_event_RT_IS16.m:228-229)
  double ax = 0.0, ad = 0.0;
  for (int k = tid; k < R; k += GN_THREADS) {
    const double a = A[k];
    if (k < R_x) ax += a; else ad += a;
  }
  ax = block_sum(ax, scratch);
  ad = block_sum(ad, scratch);
  const double A_d_mag = ad / R_d;
  double A_x_mag = ax / R_x;

  // ---- blk_sparse (src/blk_sparse.m:10-36) ----
  if (sc.blk_sparse) {
    blk_sparse_dev(Xh, Dh, st.r_blk + (size_t)slot * PL * LDF, LDF, F, l, st.rblk_pos[slot], sc, Q_s, rs1, rs2, P_s, scratch);
    if (tid == 0) st.rblk_pos[slot] = (st.rblk_pos[slot] + 1) % PL;   // every thread has read it (barriers inside)
  } else {
    for (int f = tid; f < F; f += GN_THREADS) Q_s[f] = 1.0;
  }
  __syncthreads();
  double qsum = 0.0;
  for (int f = tid; f < F; f += GN_THREADS) qsum += Q_s[f];
  qsum = block_sum(qsum, scratch);
  const double meanQ = qsum / F;

  // ---- gain (bnmf_sep_event_RT_IS16.m:221-260) ----
  double beta = 20.0 * log10(A_d_mag / A_x_mag) * sc.beta;
  if (beta < sc.beta) beta = sc.beta;
  else if (beta >= sc.beta_max) beta = sc.beta_max;
  const bool init = (l <= sc.init_N_len);
  double* ldav = st.lambda_dav + (size_t)slot * LDF;
  double* xprev = st.Xm_tilde_prev + (size_t)slot * LDF;
  for (int f = tid; f < F; f += GN_THREADS) {
    const double ym = Ym[f];
    double ld = (l == 1 && !sc.mel_mode) ? ym : ldav[f];   // Mel mode: mel_post_kernel seeded it (:205-211)
    ld = sc.alpha_d * ld + (1.0 - sc.alpha_d) * Dh[f] * beta;
    double G;
    if (sc.enhance_method == SNMFNAT_ENH_WIENER) {
      G = Xh[f] / (Xh[f] + Dh[f]);
    } else {
      double eta = (sc.alpha_eta * xprev[f] + (1.0 - sc.alpha_eta) * Xh[f] * Q_s[f]) / fmax(ld, flr);
      eta = fmax(0.0031, eta);
      G = eta / (eta + 1.0);
    }
    G = fmin(G, 1.0);
    if (init) G = 0.0 + flr;
    const double xt = G * ym;
    ldav[f] = ld;
    xprev[f] = xt;
    Xt[f] = xt;
    G_s[f] = G;
  }
  if (init) A_x_mag = flr;

  // ---- adaptation gate + history (bnmf_sep_event_RT_IS16.m:263-292) ----
  const double Q_control = (1.0 - meanQ) * sc.Ar_up;
  const bool gated = sc.adapt_train_N && (Q_control * A_d_mag > A_x_mag);
  int n_up = 0, do_up = 0;
  if (gated) {
    const int head = st.ring_head[slot];
    double* lb = st.lam_blk + ((size_t)slot * m_a + head) * LDF;
    for (int f = tid; f < F; f += GN_THREADS) {
      const double ym = Ym[f];
      lb[f] = init ? ym : ym * (f < sc.DCbin ? flr : 1.0 - G_s[f]);
    }
    double* adb = st.Ad_blk + (size_t)slot * m_a * R_a;
    for (int k = tid; k < R_a; k += GN_THREADS) adb[(size_t)head * R_a + k] = A[R_x + k];
    __syncthreads();
    const int nhead = (head + 1) % m_a;
    for (int k = tid; k < R_a; k += GN_THREADS) {
      double s = 0.0;
      bool any = false;
      for (int i = 0; i < m_a; ++i) {  // oldest column first
        const double x = adb[(size_t)((nhead + i) % m_a) * R_a + k];
        s += x;
        any |= (x != 0.0);
      }
      const int up = (Q_control * (s / m_a) > A_x_mag) ? 1 : 0;
      up_s[k] = up;
      if (up && !any) atomicExch(st.err_flag, 1);  // reference would fail with a dimension mismatch (:292 vs :323)
    }
    __syncthreads();
    if (tid == 0) {
      st.ring_head[slot] = nhead;
      if (st.update_switch[slot] == sc.update_period) {
        int nu = 0, nr = 0;
        int* iu = st.idx_up + (size_t)slot * R_a;
        int* ir = st.idx_rem + (size_t)slot * R_a;
        for (int k = 0; k < R_a; ++k) {
          if (up_s[k]) iu[nu++] = k; else ir[nr++] = k;
        }
        st.n_up[slot] = nu;
        st.do_update[slot] = (nu > 0) ? 1 : 0;
        st.update_switch[slot] = 1;
        scratch[40] = (double)nu;
        scratch[41] = (nu > 0) ? 1.0 : 0.0;
      } else {
        st.update_switch[slot] += 1;
        st.do_update[slot] = 0;
        st.n_up[slot] = 0;
        scratch[40] = 0.0;
        scratch[41] = 0.0;
      }
    }
    __syncthreads();
    n_up = (int)scratch[40];
    do_up = (int)scratch[41];
  } else if (tid == 0) {
    st.do_update[slot] = 0;
    st.n_up[slot] = 0;
  }
  if (tid == 0) {
    st.gated[slot] = gated ? 1 : 0;
    if (st.w_last && st.w_iters[slot] > 0) st.w_last[slot] = st.w_iters[slot];   // kept for the W-solve launch order
    st.w_iters[slot] = 0;
    atomicAdd(&st.stats[0], 1ull);
    atomicAdd(&st.stats[1], (unsigned long long)st.h_iters[slot]);
    if (gated) atomicAdd(&st.stats[3], 1ull);
    if (do_up) {
      atomicAdd(&st.stats[4], 1ull);
      atomicAdd(&st.stats[5], (unsigned long long)n_up);
    }
  }
  // optional copies for parity tests / the per-hop API
  for (int f = tid; f < F; f += GN_THREADS) {
    st.Q[(size_t)slot * LDF + f] = Q_s[f];
    st.G[(size_t)slot * LDF + f] = G_s[f];
  }
  if (has_trace) {
    for (int f = tid; f < F; f += GN_THREADS) {
      tr.Q[(size_t)frame * LDF + f] = Q_s[f];
      tr.G[(size_t)frame * LDF + f] = G_s[f];
    }
    for (int k = tid; k < R; k += GN_THREADS) tr.A[(size_t)frame * R + k] = A[k];
    if (tid == 0) {
      tr.info[frame * 4 + 0] = st.h_iters[slot];
      tr.info[frame * 4 + 1] = gated ? 1 : 0;
      tr.info[frame * 4 + 2] = n_up;
      tr.info[frame * 4 + 3] = 0;
    }
  }
  if (n_next >= 0) gain_order_next(d, st, g_step, n_next);
}

static size_t gain_smem_bytes(const OnlineDims& d) { return ((size_t)5 * d.F + 64) * sizeof(double) + (size_t)d.R_a * sizeof(int) + 16; }

void launch_gain(snmfnat_ctx* ctx, const OnlineDims& d, const OnlineScalars& sc, const SlotState& st,
                 const FrameArrays& fr, const TraceArrays* tr, int n_active, int g_step, int n_active_next) {
  if (n_active <= 0) return;
  const size_t smem = gain_smem_bytes(d);
  SN_REQUIRE(sc.P_len_k + sc.DCbin <= d.F && sc.P_len_k >= 2, SNMFNAT_EINVAL, "P_len_k=%d does not fit F=%d", sc.P_len_k, d.F);
  SN_REQUIRE(sc.blk_gap >= 1 && (sc.blk_gap & 1), SNMFNAT_EINVAL, "blk_gap must be odd (settings :70), got %d", sc.blk_gap);
  if (smem > 48 * 1024) SN_CUDA(cudaFuncSetAttribute(gain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  TraceArrays t{};
  if (tr) t = *tr;
  const int n_next = (st.ms_perm && st.ms_perm_step && st.ms_ticket && d.slot0 < 16 && n_active <= st.ms_perm_stride) ? n_active_next : -1;
  gain_kernel<<<dim3(n_active), dim3(GN_THREADS), smem, ctx->stream>>>(d, sc, st, fr, t, tr ? 1 : 0, g_step, n_next);
  count_launch(ctx);
  check_launch(ctx, "gain_kernel");
}

// =====================================================================================================
// W-solve (noise-basis adaptation)
// =====================================================================================================
constexpr int WS_THREADS = 256;
constexpr int WS_WARPS = WS_THREADS / 32;
constexpr int WS_CL = 4;
constexpr int WS_KT = 8;           // max k-tiles of 8 atoms  -> R_a <= 64
constexpr int WS_KMAX = WS_KT * 8;
constexpr int WS_XN = 3 * WS_KMAX + 8;

struct WsLayout {
  int NP, HSd;  // padded history length (multiple of 16), row stride of Hs (== 2 mod 8: conflict-free fragment loads)
  size_t off_H, off_red, off_xch, off_hs, off_wn, off_tot, off_Wl, off_Gl, off_scratch, bytes;
};
__host__ __device__ inline WsLayout ws_layout(int m_a) {
  WsLayout L;
  L.NP = (m_a + 15) / 16 * 16;
  L.HSd = L.NP + ((2 - L.NP % 8) + 8) % 8;
  size_t o = 0;
  L.off_H = o;       o += (size_t)WS_KMAX * L.HSd;
  L.off_red = o;     o += (size_t)2 * WS_WARPS * WS_KMAX;
  L.off_xch = o;     o += (size_t)2 * WS_XN;
  L.off_hs = o;      o += WS_KMAX;
  L.off_wn = o;      o += WS_KMAX;
  L.off_tot = o;     o += 2 * WS_KMAX;
  L.off_Wl = o;      o += 8 * WS_KMAX;   // up to 7 leftover rows
  L.off_Gl = o;      o += 8 * WS_KMAX;
  L.off_scratch = o; o += 64;
  L.bytes = o * sizeof(double);
  return L;
}
size_t wsolve_smem_bytes(const OnlineDims& d) { return ws_layout(d.m_a).bytes; }

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
// sum over the 8 lanes that share lane&3 (the 8 rows of a fragment)
__device__ __forceinline__ double rows8_sum(double v) {
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  return v;
}

__global__ void __cluster_dims__(WS_CL, 1, 1) __launch_bounds__(WS_THREADS, 1)
wsolve_kernel(OnlineDims d, OnlineScalars sc, SlotState st, TraceArrays tr, int has_trace, int g_step) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int slot = d.slot0 + (int)(blockIdx.x / WS_CL) * d.slot_stride;
  const int l = g_step + 1 - st.l_offset[slot];
  if (l < 1 || l > st.n_hops[slot]) return;
  if (!st.do_update[slot]) return;  // uniform over the cluster

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, tg = lane & 3;
  const int F = d.F, LDF = d.LDF, R_a = d.R_a, R_d = d.R_d, n = d.m_a;
  const double flr = sc.flr;
  const WsLayout L = ws_layout(n);
  const int HSd = L.HSd, NP = L.NP;

  extern __shared__ __align__(16) double smem[];
  double* Hs = smem + L.off_H;
  double* red = smem + L.off_red;      // [2][WS_WARPS][WS_KMAX]
  double* xch = smem + L.off_xch;      // [2][WS_XN]
  double* hs_s = smem + L.off_hs;      // [KMAX] sum_t H
  double* wn_s = smem + L.off_wn;
  double* tot = smem + L.off_tot;      // [2][KMAX] combined reductions
  double* Wl = smem + L.off_Wl;        // leftover rows [row][KMAX]
  double* Gl = smem + L.off_Gl;
  double* scratch = smem + L.off_scratch;

  const int Ru = st.n_up[slot];
  const int kt = (Ru + 7) / 8;
  const int* __restrict__ idx_up = st.idx_up + (size_t)slot * R_a;
  const int* __restrict__ idx_rem = st.idx_rem + (size_t)slot * R_a;
  const int sel = st.bd_sel[slot];
  const double* __restrict__ Bcur = st.Bd[sel] + (size_t)slot * R_d * LDF;
  double* __restrict__ Bnext = st.Bd[sel ^ 1] + (size_t)slot * R_d * LDF;
  const double* __restrict__ Vg = st.lam_blk + (size_t)slot * n * LDF;
  const double* __restrict__ Adb = st.Ad_blk + (size_t)slot * n * R_a;

  // row tiles: NFT full tiles of 8 rows, two per warp; leftover rows (F % 8) on the last warp of the cluster
  const int NFT = F / 8;
  const int gw = rank * WS_WARPS + warp;
  const int nleft = F - NFT * 8;
  const bool has_left = (nleft > 0) && (gw == WS_CL * WS_WARPS - 1);
  int f0t[2];
  bool tv[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int ti = 2 * gw + t;
    tv[t] = ti < NFT;
    f0t[t] = ti * 8;
  }

  // ---- load W tiles (fragment order: w[t][j][e] = W[f0+g][8j+2tg+e]) and column sums of squares ----
  double w[2][WS_KT][2], gacc[2][WS_KT][2];
#pragma unroll
  for (int t = 0; t < 2; ++t)
#pragma unroll
    for (int j = 0; j < WS_KT; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int k = 8 * j + 2 * tg + e;
        double x = 0.0;
        if (tv[t] && k < Ru) x = Bcur[(size_t)idx_up[k] * LDF + f0t[t] + g];
        w[t][j][e] = x;
        gacc[t][j][e] = 0.0;
      }
  if (has_left)
    for (int i = lane; i < nleft * WS_KMAX; i += 32) {
      const int r = i / WS_KMAX, k = i % WS_KMAX;
      Wl[i] = (k < Ru) ? Bcur[(size_t)idx_up[k] * LDF + NFT * 8 + r] : 0.0;
    }
  __syncwarp();

  // per-warp partial of a per-column quantity -> red[which][warp][k]
  auto warp_partial = [&](int which, auto&& f_tile, auto&& f_left) {
    double* rw = red + ((size_t)which * WS_WARPS + warp) * WS_KMAX;
#pragma unroll
    for (int j = 0; j < WS_KT; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        double s = 0.0;
#pragma unroll
        for (int t = 0; t < 2; ++t) s += f_tile(t, j, e);
        s = rows8_sum(s);
        if (g == 0) rw[8 * j + 2 * tg + e] = s;
      }
    __syncwarp();
    if (has_left)
      for (int k = lane; k < WS_KMAX; k += 32) {
        double s = rw[k];
        for (int r = 0; r < nleft; ++r) s += f_left(r, k);
        rw[k] = s;
      }
  };
  // CTA partial (fixed warp order) -> exchange buffer; cluster barrier; totals in rank order -> tot[which][k]
  auto cluster_combine = [&](int nwhich, int xbuf, double extra, double* extra_out) {
    __syncthreads();
    double* xb = xch + (size_t)xbuf * WS_XN;
    for (int i = tid; i < nwhich * WS_KMAX; i += WS_THREADS) {
      const int which = i / WS_KMAX, k = i % WS_KMAX;
      double s = 0.0;
      for (int ww = 0; ww < WS_WARPS; ++ww) s += red[((size_t)which * WS_WARPS + ww) * WS_KMAX + k];
      xb[which * WS_KMAX + k] = s;
    }
    if (tid == 0) xb[3 * WS_KMAX] = extra;
    cluster.sync();
    for (int i = tid; i < nwhich * WS_KMAX; i += WS_THREADS) {
      double s = 0.0;
      for (int c = 0; c < WS_CL; ++c) s += cluster.map_shared_rank(xb, c)[i];
      tot[i] = s;
    }
    if (extra_out) {
      double s = 0.0;
      for (int c = 0; c < WS_CL; ++c) s += cluster.map_shared_rank(xb, c)[3 * WS_KMAX];
      *extra_out = s;
    }
    __syncthreads();
  };

  // column norms of init_w (sparse_nmf.m:158)
  warp_partial(0, [&](int t, int j, int e) { return w[t][j][e] * w[t][j][e]; },
               [&](int r, int k) { return Wl[r * WS_KMAX + k] * Wl[r * WS_KMAX + k]; });
  cluster_combine(1, 1, 0.0, nullptr);
  if (tid < WS_KMAX) wn_s[tid] = (tid < Ru) ? sqrt(tot[tid]) : 1.0;
  __syncthreads();
#pragma unroll
  for (int t = 0; t < 2; ++t)
#pragma unroll
    for (int j = 0; j < WS_KT; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) w[t][j][e] = w[t][j][e] / wn_s[8 * j + 2 * tg + e];  // :159
  if (has_left)
    for (int i = lane; i < nleft * WS_KMAX; i += 32) Wl[i] = Wl[i] / wn_s[i % WS_KMAX];
  // H = init_h .* wn (:160), zero padded; row sums (constant over the solve)
  for (int i = tid; i < WS_KMAX * NP; i += WS_THREADS) {
    const int k = i / NP, t = i % NP;
    double x = 0.0;
    if (k < Ru && t < n) x = Adb[(size_t)t * R_a + idx_up[k]] * wn_s[k];
    Hs[(size_t)k * HSd + t] = x;
  }
  __syncthreads();
  if (tid < WS_KMAX) {
    double s = 0.0;
    for (int t = 0; t < n; ++t) s += Hs[(size_t)tid * HSd + t];
    hs_s[tid] = s;
  }
  __syncthreads();
  double hsum_all = 0.0;
  for (int k = 0; k < Ru; ++k) hsum_all += hs_s[k];

  // ---- multiplicative updates ----
  int it = 0;
  double last_cost = INFINITY, cost = 0.0;
  const int ngroups = NP / 16;
  for (;;) {
    double cacc = 0.0;
    for (int tgp = 0; tgp < ngroups; ++tgp) {
      const int n0 = tgp * 16;
      // GEMM 1: lambda tile = W * H for 16 history columns (even columns -> c[.][0], odd -> c[.][1])
      double c[2][2][2];
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int eo = 0; eo < 2; ++eo) c[t][eo][0] = c[t][eo][1] = 0.0;
#pragma unroll
      for (int j = 0; j < WS_KT; ++j)
        if (j < kt) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const double2 b = *reinterpret_cast<const double2*>(Hs + (size_t)(8 * j + 2 * tg + e) * HSd + n0 + 2 * g);
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              dmma(c[t][0][0], c[t][0][1], w[t][j][e], b.x);
              dmma(c[t][1][0], c[t][1][1], w[t][j][e], b.y);
            }
          }
        }
      // ratio v./lambda and cost terms; this lane holds history columns n0+4tg+q, q=0..3
      double rt[2][4];
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int tcol = n0 + 4 * tg + q;
          double r = 0.0;
          if (tv[t] && tcol < n) {
            const double lam = fmax(c[t][q & 1][q >> 1], flr);                          // :243
            const double v = fmax(Vg[(size_t)tcol * LDF + f0t[t] + g], flr);           // :169
            r = v / lam;
            if (sc.cost_check && it >= 1) cacc += v * log(r) - v + lam;                 // :250
          }
          rt[t][q] = r;
        }
      // GEMM 2: G tile += (v./lambda) * H'
#pragma unroll
      for (int j = 0; j < WS_KT; ++j)
        if (j < kt) {
          const double* hp = Hs + (size_t)(8 * j + g) * HSd + n0 + 4 * tg;
          const double2 h01 = *reinterpret_cast<const double2*>(hp);
          const double2 h23 = *reinterpret_cast<const double2*>(hp + 2);
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            dmma(gacc[t][j][0], gacc[t][j][1], rt[t][0], h01.x);
            dmma(gacc[t][j][0], gacc[t][j][1], rt[t][1], h01.y);
            dmma(gacc[t][j][0], gacc[t][j][1], rt[t][2], h23.x);
            dmma(gacc[t][j][0], gacc[t][j][1], rt[t][3], h23.y);
          }
        }
    }
    // leftover rows: lane-parallel over history columns
    if (has_left) {
      for (int r = 0; r < nleft; ++r) {
        const int f = NFT * 8 + r;
        double rl[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int tcol = lane + 32 * q;
          double rr = 0.0;
          if (tcol < n) {
            double lam = 0.0;
            for (int k = 0; k < Ru; ++k) lam = fma(Wl[r * WS_KMAX + k], Hs[(size_t)k * HSd + tcol], lam);
            lam = fmax(lam, flr);
            const double v = fmax(Vg[(size_t)tcol * LDF + f], flr);
            rr = v / lam;
            if (sc.cost_check && it >= 1) cacc += v * log(rr) - v + lam;
          }
          rl[q] = rr;
        }
        for (int k = 0; k < Ru; ++k) {
          double s = 0.0;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (lane + 32 * q < NP) s = fma(rl[q], Hs[(size_t)k * HSd + lane + 32 * q], s);
          s = warp_sum(s);
          if (lane == 0) Gl[r * WS_KMAX + k] = s;
        }
      }
      __syncwarp();
    }
    // column reductions: cw_k = sum_f w, s_k = sum_f G.*w                               :215-221
    warp_partial(0, [&](int t, int j, int e) { return w[t][j][e]; },
                 [&](int r, int k) { return (k < Ru) ? Wl[r * WS_KMAX + k] : 0.0; });
    warp_partial(1, [&](int t, int j, int e) { return gacc[t][j][e] * w[t][j][e]; },
                 [&](int r, int k) { return (k < Ru) ? Gl[r * WS_KMAX + k] * Wl[r * WS_KMAX + k] : 0.0; });
    const double cpart = block_sum(cacc, scratch);
    double div = 0.0;
    cluster_combine(2, 0, cpart, &div);
    bool stop = false;
    if (sc.cost_check && it >= 1) {
      cost = div + sc.sparsity * hsum_all;                                               // :261
      if (it > 1 && sc.conv_eps > 0.0) {
        const double e = fabs(cost - last_cost) / last_cost;
        if (e < sc.conv_eps) stop = true;
      }
      last_cost = cost;
    }
    if (it >= sc.max_iter) stop = true;
    if (stop) break;
    // W update                                                                          :215-222
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int j = 0; j < WS_KT; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int k = 8 * j + 2 * tg + e;
          const double wv = w[t][j][e];
          const double hs = hs_s[k];
          const double dpw = fmax(hs + tot[WS_KMAX + k] * wv, flr);
          const double dmw = gacc[t][j][e] + (hs * tot[k]) * wv;
          w[t][j][e] = (k < Ru) ? wv * dmw / dpw : 0.0;
          gacc[t][j][e] = 0.0;
        }
    if (has_left)
      for (int i = lane; i < nleft * WS_KMAX; i += 32) {
        const int k = i % WS_KMAX;
        const double wv = Wl[i];
        const double hs = hs_s[k];
        const double dpw = fmax(hs + tot[WS_KMAX + k] * wv, flr);
        const double dmw = Gl[i] + (hs * tot[k]) * wv;
        Wl[i] = (k < Ru) ? wv * dmw / dpw : 0.0;
      }
    __syncwarp();
    // column normalisation                                                              :242
    warp_partial(0, [&](int t, int j, int e) { return w[t][j][e] * w[t][j][e]; },
                 [&](int r, int k) { return Wl[r * WS_KMAX + k] * Wl[r * WS_KMAX + k]; });
    cluster_combine(1, 1, 0.0, nullptr);
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int j = 0; j < WS_KT; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int k = 8 * j + 2 * tg + e;
          if (k < Ru) w[t][j][e] = w[t][j][e] / sqrt(tot[k]);
        }
    if (has_left)
      for (int i = lane; i < nleft * WS_KMAX; i += 32) {
        const int k = i % WS_KMAX;
        if (k < Ru) Wl[i] = Wl[i] / sqrt(tot[k]);
      }
    __syncwarp();
    ++it;
  }

  // ---- B_DFT_d = [B_rem, B_new, B_fix]  (bnmf_sep_event_RT_IS16.m:336) into the other buffer ----
  const int n_rem = R_a - Ru;
#pragma unroll
  for (int t = 0; t < 2; ++t)
#pragma unroll
    for (int j = 0; j < WS_KT; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int k = 8 * j + 2 * tg + e;
        if (tv[t] && k < Ru) Bnext[(size_t)(n_rem + k) * LDF + f0t[t] + g] = w[t][j][e];
      }
  if (has_left)
    for (int i = lane; i < nleft * WS_KMAX; i += 32) {
      const int r = i / WS_KMAX, k = i % WS_KMAX;
      if (k < Ru) Bnext[(size_t)(n_rem + k) * LDF + NFT * 8 + r] = Wl[i];
    }
  {
    const double* __restrict__ Bfix = st.Bd_fix + (size_t)slot * st.bdfix_stride;
    const int rb = F / WS_CL, rr = F % WS_CL;
    const int rows = rb + (rank < rr ? 1 : 0);
    const int r0 = rank * rb + (rank < rr ? rank : rr);
    const int ncopy = n_rem + (R_d - R_a);
    for (int i = tid; i < ncopy * rows; i += WS_THREADS) {
      const int cidx = i / rows, f = r0 + i % rows;
      if (cidx < n_rem) Bnext[(size_t)cidx * LDF + f] = Bcur[(size_t)idx_rem[cidx] * LDF + f];
      else {
        const int k = R_a + (cidx - n_rem);
        Bnext[(size_t)k * LDF + f] = Bfix[(size_t)k * LDF + f];
      }
    }
  }
  if (rank == 0 && tid == 0) {
    st.w_iters[slot] = it;
    atomicAdd(&st.stats[2], (unsigned long long)it);
    if (has_trace) tr.info[(st.frame_base[slot] + g_step) * 4 + 3] = it;
  }
  cluster.sync();  // peers may still be reading this CTA's exchange buffers
  if (rank == 0 && tid == 0) st.bd_sel[slot] = sel ^ 1;
}

void launch_wsolve(snmfnat_ctx* ctx, const OnlineDims& d, const OnlineScalars& sc, const SlotState& st,
                   const TraceArrays* tr, int n_active, int g_step) {
  if (n_active <= 0 || !sc.adapt_train_N) return;
  if (!force_generic() && wsolve_fast_supported(ctx, d)) {
    launch_wsolve_fast(ctx, d, sc, st, tr, n_active, g_step);
    return;
  }
  SN_REQUIRE(d.R_a <= WS_KMAX, SNMFNAT_EUNSUPPORTED, "W-solve kernel supports R_a <= %d (got %d)", WS_KMAX, d.R_a);
  SN_REQUIRE(d.F / 8 <= 2 * WS_CL * WS_WARPS, SNMFNAT_EUNSUPPORTED, "W-solve kernel supports F <= %d (got %d)",
             16 * WS_CL * WS_WARPS + 7, d.F);
  SN_REQUIRE(d.m_a <= 128, SNMFNAT_EUNSUPPORTED, "W-solve kernel supports m_a <= 128 (got %d)", d.m_a);
  const size_t smem = ws_layout(d.m_a).bytes;
  SN_REQUIRE((int)smem <= ctx->max_smem_optin, SNMFNAT_EUNSUPPORTED, "W-solve shared memory %zu too large", smem);
  SN_CUDA(cudaFuncSetAttribute(wsolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  TraceArrays t{};
  if (tr) t = *tr;
  wsolve_kernel<<<dim3(WS_CL * n_active), dim3(WS_THREADS), smem, ctx->stream>>>(d, sc, st, t, tr ? 1 : 0, g_step);
  count_launch(ctx);
  check_launch(ctx, "wsolve_kernel");
}

}  // namespace snmfnat
