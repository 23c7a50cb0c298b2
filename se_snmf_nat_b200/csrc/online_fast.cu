// Second-generation per-hop MU kernels for the shipped geometry (fftlength 1024 -> F = 513 rows).
//
// hsolve_fast_kernel : 4-CTA cluster per stream, 16 warps per CTA.  Each CTA keeps 128 rows of W = [B_x B_d] in
//   shared memory with an XOR swizzle, element (k,f) at k*128 + (f ^ 2*(k & 7)) (row pairs permuted), so that BOTH directions of the
//   mat-vec pair are bank-conflict free without any cross-lane reduction:
//     lambda = W h   : lanes <-> rows   (one column k per step, h_k broadcast)
//     g = W'(v./lambda): lanes <-> atoms (one row pair per step, ratio broadcast)
//   The F - 512 tail rows (the Nyquist bin for F = 513) live in a small row-major side array on the last rank.
//   Per MU iteration: 3 block barriers + 1 cluster barrier; the R-vector g, the cost partial and the tail
//   terms travel through distributed shared memory.
//
// wsolve_fast_kernel : 4-CTA cluster per stream, 16 warps per CTA, one 8-row FP64 tensor-core tile per warp
//   (mma.sync m8n8k4.f64); the F % 8 leftover rows are plain FMAs shared out over the warps of the last CTA at the top of
//   every pass (a whole tensor-core tile for one row would give one scheduler of that SM a fifth tile; a 17th warp for
//   them caps every thread at 96 registers).  W and G = (V./Lambda) H' fragments stay in registers for the whole solve in the SAME
//   fragment layout (the k order of the first GEMM is permuted to match the accumulator layout of the second), the
//   CTA's slice of V = lambda_d_blk is staged once in shared memory, H is staged once.  Padding rows/columns are
//   neutral by construction (V pad = floor, W pad = 0, H pad = 0) so the inner loops carry no predicates.
//   v./lambda uses a Newton reciprocal and the KL cost a table-driven log (|err| < 4e-16 absolute).
#include <cooperative_groups.h>
#include <type_traits>
#include <cmath>
#include "online.cuh"
#include "online_dev.cuh"

namespace cg = cooperative_groups;

namespace snmfnat {

// =====================================================================================================
// shared helpers
// =====================================================================================================
__global__ void log_table_kernel(double2* tab) {
  const int i = threadIdx.x;
  if (i < 128) {
    const double c = 1.0 + (i + 0.5) / 128.0;
    tab[i] = make_double2(1.0 / c, log(c));
  }
}

static double2* g_log_tab[64] = {nullptr};
const double2* log_table(snmfnat_ctx* ctx) {
  const int dev = ctx->device;
  SN_REQUIRE(dev >= 0 && dev < 64, SNMFNAT_EINVAL, "device index out of range");
  if (!g_log_tab[dev]) {
    double2* p = nullptr;
    SN_CUDA(cudaMalloc(&p, 128 * sizeof(double2)));
    log_table_kernel<<<1, 128, 0, ctx->stream>>>(p);
    count_launch(ctx);
    SN_CUDA(cudaStreamSynchronize(ctx->stream));
    g_log_tab[dev] = p;
  }
  return g_log_tab[dev];
}

// =====================================================================================================
// H-solve, fast path: F = 512 + E (0 <= E <= 8)
// =====================================================================================================
constexpr int HF_THREADS = 512;
constexpr int HF_WARPS = 16;
constexpr int HF_CL = 4;
constexpr int HF_ROWS = 128;   // rows per CTA
constexpr int HF_KG = 8;       // k groups in the lambda pass

constexpr int HF_HP = 34;      // stride of the per-k-group copy of h (>= ceil(R/8), even)
struct HfLayout {
  int E, XN;
  size_t off_W, off_Wt, off_v, off_r, off_lam, off_h, off_hp, off_dph, off_recv, off_misc, off_bar, bytes;
};
__host__ __device__ inline HfLayout hf_layout(int F, int R) {
  HfLayout L;
  L.E = F - HF_CL * HF_ROWS;
  L.XN = R + 2;
  size_t o = 0;
  L.off_W = o;    o += (size_t)R * HF_ROWS;
  L.off_Wt = o;   o += (size_t)(L.E > 0 ? L.E : 0) * R;
  o = (o + 1) & ~(size_t)1;
  L.off_v = o;    o += HF_ROWS + 8;
  L.off_r = o;    o += HF_ROWS + 8;
  L.off_lam = o;  o += (size_t)(HF_KG / 2) * HF_ROWS;   // 4 partial groups: the two k groups of a warp are added by shuffle
  L.off_h = o;    o += R;
  o = (o + 1) & ~(size_t)1;
  L.off_hp = o;   o += (size_t)HF_KG * HF_HP;
  L.off_dph = o;  o += R;
  o = (o + 1) & ~(size_t)1;
  // [2][4][XN]: pushed partials of g, [R] cost partial, [R+1] sum(h).  Its head doubles as the [2][XN] buffer of the
  // one-off normalisation exchange before the loop.
  L.off_recv = o; o += 2 * (size_t)HF_CL * L.XN;
  L.off_misc = o; o += 48;
  L.off_bar = o;  o += 2;                          // 2 mbarriers
  L.bytes = o * sizeof(double);
  return L;
}

__global__ void __cluster_dims__(HF_CL, 1, 1) __launch_bounds__(HF_THREADS, 1)
hsolve_fast_kernel(OnlineDims d, OnlineScalars sc, SlotState st, FrameArrays fr, const double* __restrict__ h_init,
                   int g_step, const double2* __restrict__ log_tab) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int slot = d.slot0 + (int)(blockIdx.x / HF_CL) * d.slot_stride;
  const int l = g_step + 1 - st.l_offset[slot];
  if (l < 1 || l > st.n_hops[slot]) return;  // uniform over the cluster

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int F = d.F, R = d.R, R1 = d.R_x, LDF = d.LDF;
  const HfLayout L = hf_layout(F, R);
  const int E = L.E;
  const bool tail_rank = (rank == HF_CL - 1) && E > 0;
  const int f0 = rank * HF_ROWS;

  extern __shared__ __align__(1024) double smem[];
  double* Ws = smem + L.off_W;
  double* Wt = smem + L.off_Wt;       // [E][R] tail rows (last rank only)
  double* v_s = smem + L.off_v;       // [128 + E]
  double* r_s = smem + L.off_r;       // ratio v./lambda
  double* lam_part = smem + L.off_lam;
  double* h_s = smem + L.off_h;
  double* hp_s = smem + L.off_hp;     // [8][HF_HP]: h of the atoms k = kg + 8 j, grouped by kg (lambda pass reads pairs)
  double* dph_s = smem + L.off_dph;
  double* xch = smem + L.off_recv;    // [2][XN] of the normalisation exchange (aliases the head of recv; see below)
  double* recv = smem + L.off_recv;   // [2][4][XN] partials pushed by the 4 CTAs of the cluster
  double* misc = smem + L.off_misc;   // [0..7] cost partials per warp, [8] hsum, [16..23] lambda of tail rows
  const unsigned bar0 = (unsigned)__cvta_generic_to_shared(smem + L.off_bar);
  if (tid == 0) {
    hf_mbar_init(bar0, 1);
    hf_mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  const double* __restrict__ W1 = st.Bx;
  const double* __restrict__ W2 = st.Bd[st.bd_sel[slot]] + (size_t)slot * d.R_d * LDF;
  const long long frame = st.frame_base[slot] + g_step;
  const double* __restrict__ V = fr.Ym + (size_t)frame * LDF;
  const double flr = sc.flr;

  // ---- stage W (swizzled), partial column sums / sums of squares over this CTA's rows.  Four columns per step: their 16
  //      (+ tail) loads are in flight together, the private half of the basis comes from HBM ----
  constexpr int SC = 4;
  for (int k0 = warp; k0 < R; k0 += SC * HF_WARPS) {
    double x[SC][HF_ROWS / 32], xt[SC];
#pragma unroll
    for (int c = 0; c < SC; ++c) {
      const int k = k0 + c * HF_WARPS;
      const double* src = (k < R1 ? W1 + (size_t)k * LDF : W2 + (size_t)(k - R1) * LDF);
#pragma unroll
      for (int j = 0; j < HF_ROWS / 32; ++j) x[c][j] = (k < R) ? src[f0 + lane + 32 * j] : 0.0;
      xt[c] = (k < R && tail_rank && lane < E) ? src[HF_CL * HF_ROWS + lane] : 0.0;
    }
#pragma unroll
    for (int c = 0; c < SC; ++c) {
      const int k = k0 + c * HF_WARPS;
      if (k >= R) break;
      double s1 = 0.0, s2 = 0.0;
      const int sw = (k & 7) << 1;
#pragma unroll
      for (int j = 0; j < HF_ROWS / 32; ++j) {
        const int f = lane + 32 * j;
        Ws[(size_t)k * HF_ROWS + (f ^ sw)] = x[c][j];
        s1 += x[c][j];
        s2 = fma(x[c][j], x[c][j], s2);
      }
      if (tail_rank && lane < E) {
        Wt[(size_t)lane * R + k] = xt[c];
        s1 += xt[c];
        s2 = fma(xt[c], xt[c], s2);
      }
      s1 = warp_sum(s1);
      s2 = warp_sum(s2);
      if (lane == 0) {
        xch[k] = s2;  // buffer 0 carries [sumsq | sum] for the normalisation exchange (R + 8 >= ... uses both buffers)
        xch[L.XN + k] = s1;
      }
    }
  }
  for (int f = tid; f < HF_ROWS + (tail_rank ? E : 0); f += HF_THREADS)
    v_s[f] = fmax(V[f < HF_ROWS ? f0 + f : HF_CL * HF_ROWS + (f - HF_ROWS)], flr);  // sparse_nmf.m:169
  cluster.sync();

  // ---- column norms, h scaling (sparse_nmf.m:157-160), denominators (:192-193) ----
  double wn_r = 1.0, dph_r = 0.0;   // of atom `tid`
  if (tid < R) {
    double s2 = 0.0, s1 = 0.0;
    for (int c = 0; c < HF_CL; ++c) {
      const double* rx = cluster.map_shared_rank(xch, c);
      s2 += rx[tid];
      s1 += rx[L.XN + tid];
    }
    const double wn = sqrt(s2);
    wn_r = wn;
    dph_s[tid] = 1.0 / wn;                                  // staging of 1/wn for the scaling loop below
    dph_r = 1.0 / fmax(s1 / wn + sc.sparsity, flr);         // reciprocal of the H-update denominator (:192-193)
    h_s[tid] = h_init[tid] * wn;
    hp_s[(tid & 7) * HF_HP + (tid >> 3)] = h_s[tid];
  }
  cluster.sync();  // everyone has read both exchange buffers before they are reused; publishes wn_s / h_s
  for (int k = warp; k < R; k += HF_WARPS) {
    const double inv = dph_s[k];        // one division per column (above); the elements are scaled by the reciprocal
#pragma unroll
    for (int j = 0; j < HF_ROWS / 32; ++j) {
      double* p = Ws + (size_t)k * HF_ROWS + lane + 32 * j;
      *p = *p * inv;
    }
    if (tail_rank && lane < E) Wt[(size_t)lane * R + k] = Wt[(size_t)lane * R + k] * inv;
  }
  __syncthreads();

  // ---- multiplicative updates ----
  // lambda pass roles: warp = (pair of k groups kp / kp+4, 32-row quarter rq); the half-warps take one k group each
  const int kp = warp & 3, rq = warp >> 2, half = lane >> 4, l16 = lane & 15;
  const int kg = kp + 4 * half;
  const int kb = warp;                                  // g pass: block of 16 atoms, lanes 16..31 take rows 64..127
  const int kl = lane & 15, fh = lane >> 4;
  const int kB = kb * 16 + kl;
  // lambda partials of W x for the vector staged in hp_s (grouped by k & 7) / h_s: every lane sums the columns
  // k = kg, kg+8, ... for the row pair rq*32 + 2*l16 + {0,1} (one 16-byte load per column; the x of two consecutive
  // columns of the group comes as one pair); the two half-warps are added by shuffle -> 4 partial groups per row
  auto lambda_pass = [&]() {
    double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
    const double* base = Ws + rq * 32 + ((2 * l16) ^ (kg << 1));   // k & 7 == kg for every column of this group
    const double* hp = hp_s + kg * HF_HP;
    int k = kg, j = 0;
#pragma unroll 4
    for (; k + 8 < R; k += 16, j += 2) {
      const double2 hh = *reinterpret_cast<const double2*>(hp + j);
      const double2 w0 = *reinterpret_cast<const double2*>(base + (size_t)k * HF_ROWS);
      const double2 w1 = *reinterpret_cast<const double2*>(base + (size_t)(k + 8) * HF_ROWS);
      a0 = fma(w0.x, hh.x, a0);
      a1 = fma(w0.y, hh.x, a1);
      b0 = fma(w1.x, hh.y, b0);
      b1 = fma(w1.y, hh.y, b1);
    }
    if (k < R) {
      const double h0 = hp[j];
      const double2 w0 = *reinterpret_cast<const double2*>(base + (size_t)k * HF_ROWS);
      a0 = fma(w0.x, h0, a0);
      a1 = fma(w0.y, h0, a1);
    }
    a0 += b0;
    a1 += b1;
    a0 += __shfl_xor_sync(0xffffffffu, a0, 16);
    a1 += __shfl_xor_sync(0xffffffffu, a1, 16);
    if (half == 0) *reinterpret_cast<double2*>(lam_part + kp * HF_ROWS + rq * 32 + 2 * l16) = make_double2(a0, a1);
    if (tail_rank && warp < E) {  // tail row `warp`: lanes over atoms
      double s = 0.0;
      for (int kk = lane; kk < R; kk += 32) s = fma(Wt[(size_t)warp * R + kk], h_s[kk], s);
      s = warp_sum(s);
      if (lane == 0) misc[16 + warp] = s;
    }
  };
  int it = 0, buf = 0;
  double last_cost = INFINITY, cost = 0.0;
  for (;;) {
    lambda_pass();   // (A)
    __syncthreads();
    // (R) ratio + cost terms for the CTA's rows (threads 0..127) and the tail rows (threads 128..128+E)
    if (warp < 5) {
      const bool main_row = tid < HF_ROWS;
      const bool tail_row = tail_rank && tid >= HF_ROWS && tid < HF_ROWS + E;
      if (main_row || tail_row) {
        double lam = 0.0;
        if (main_row) {
#pragma unroll
          for (int q = 0; q < HF_KG / 2; ++q) lam += lam_part[q * HF_ROWS + tid];
        } else {
          lam = misc[16 + tid - HF_ROWS];
        }
        lam = fmax(lam, flr);
        r_s[tid] = v_s[tid] * fast_rcp(lam);
        // lambda is kept for the cost terms, which a spare warp evaluates during phase (B), off the critical path
        if (main_row) lam_part[tid] = lam; else misc[16 + tid - HF_ROWS] = lam;
      }
    }
    __syncthreads();
    // (B) g partial over this CTA's rows: lane <-> atom kB, half-warps split the rows, pairs of rows per step.  Every
    //     partial is PUSHED to the 4 CTAs of the cluster by the lane that produced it (st.async; the receiver's
    //     mbarrier counts the bytes): no cluster barrier, no fence, no remote loads on the critical path.
    const unsigned bar = bar0 + 8u * (unsigned)buf;
    const unsigned parity = (unsigned)(it >> 1) & 1u;
    double* rb = recv + (size_t)buf * HF_CL * L.XN;
    const unsigned my_row = (unsigned)__cvta_generic_to_shared(rb + (size_t)rank * L.XN);
    if (tid == 0) hf_mbar_expect_tx(bar, (unsigned)((HF_CL * (R + 1) + 1) * sizeof(double)));
    if (kb * 16 < R) {
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      if (kB < R) {
        // row pair i of atom kB lives at pair slot i ^ (kB & 7); the row base is 512-byte aligned, so the byte address
        // of that slot is (base ^ ((kB & 7) << 4)) ^ (i << 4): one LOP3 per 16-byte load
        const unsigned wx = (unsigned)__cvta_generic_to_shared(Ws + (size_t)kB * HF_ROWS + fh * 64) ^ ((unsigned)(kB & 7) << 4);
        const double2* rp = reinterpret_cast<const double2*>(r_s + fh * 64);
#pragma unroll 8
        for (int i = 0; i < 32; i += 2) {
          double2 w0, w1;
          asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(w0.x), "=d"(w0.y) : "r"(wx ^ ((unsigned)i << 4)));
          asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(w1.x), "=d"(w1.y) : "r"(wx ^ ((unsigned)(i + 1) << 4)));
          const double2 q0 = rp[i], q1 = rp[i + 1];
          a0 = fma(w0.x, q0.x, a0);
          a1 = fma(w0.y, q0.y, a1);
          a2 = fma(w1.x, q1.x, a2);
          a3 = fma(w1.y, q1.y, a3);
        }
      }
      double g = (a0 + a1) + (a2 + a3);
      g += __shfl_xor_sync(0xffffffffu, g, 16);
      if (fh == 0 && kB < R) {
        if (tail_rank)
          for (int e = 0; e < E; ++e) g = fma(Wt[(size_t)e * R + kB], r_s[HF_ROWS + e], g);
        const unsigned la = my_row + 8u * (unsigned)kB;
#pragma unroll
        for (int c = 0; c < HF_CL; ++c) hf_st_async(hf_mapa(la, c), g, hf_mapa(bar, c));
      }
    } else if (warp == HF_WARPS - 1) {
      double s = 0.0;  // sum of h for the sparsity term of the cost (identical order on every CTA); stays in this CTA
      for (int kk = lane; kk < R; kk += 32) s += h_s[kk];
      s = warp_sum(s);
      if (lane == 0) hf_st_async(hf_mapa(my_row + 8u * (unsigned)(R + 1), rank), s, hf_mapa(bar, rank));
    } else if (warp == HF_WARPS - 2) {
      double cterm = 0.0;  // KL divergence terms of this CTA's rows                  sparse_nmf.m:250
      if (sc.cost_check && it >= 1) {
#pragma unroll
        for (int j = 0; j < HF_ROWS / 32; ++j) {
          const int f = lane + 32 * j;
          const double v = v_s[f], lam = lam_part[f];
          cterm += fma(v, fast_log(r_s[f], log_tab), lam - v);
        }
        if (tail_rank && lane < E) {
          const double v = v_s[HF_ROWS + lane], lam = misc[16 + lane];
          cterm += fma(v, fast_log(r_s[HF_ROWS + lane], log_tab), lam - v);
        }
      }
      cterm = warp_sum(cterm);
      if (lane == 0) {
        const unsigned la = my_row + 8u * (unsigned)R;
#pragma unroll
        for (int c = 0; c < HF_CL; ++c) hf_st_async(hf_mapa(la, c), cterm, hf_mapa(bar, c));
      }
    }
    hf_mbar_wait(bar, parity);
    // (C) combine the 4 CTAs in rank order, convergence test, h update
    double gk = 0.0;
    if (tid < R)
      for (int c = 0; c < HF_CL; ++c) gk += rb[(size_t)c * L.XN + tid];
    bool stop = false;
    if (sc.cost_check && it >= 1) {
      double div = 0.0;
      for (int c = 0; c < HF_CL; ++c) div += rb[(size_t)c * L.XN + R];
      cost = div + sc.sparsity * rb[(size_t)rank * L.XN + R + 1];          // :261
      if (it > 1 && sc.conv_eps > 0.0) {
        const double e = fabs(cost - last_cost) / last_cost;             // :274
        if (e < sc.conv_eps) stop = true;
      }
      last_cost = cost;
    }
    if (it >= sc.max_iter) stop = true;
    if (stop) break;
    if (tid < R) {
      const double hn = h_s[tid] * gk * dph_r;                             // :195
      h_s[tid] = hn;
      hp_s[(tid & 7) * HF_HP + (tid >> 3)] = hn;
    }
    __syncthreads();
    buf ^= 1;
    ++it;
  }

  // ---- outputs ----
  __syncthreads();
  double h_fin = 0.0;
  if (tid < R) {
    h_fin = h_s[tid];
    if (rank == 0) st.A[(size_t)slot * R + tid] = h_fin;
  }
  if (rank == 0 && tid == 0) {
    st.h_iters[slot] = it;
    st.h_cost[slot] = cost;
  }
  // Both reconstructions in ONE pass over the resident basis: X_hat = B_x A_x sums the columns k < R1, D_hat = B_d A_d the
  // others (bnmf_sep_event_RT_IS16.m:174,197).  Each column feeds the accumulators of its own class only, in the same
  // order as a pass with the other class zeroed would, so the sums are bit-identical to two separate passes.
  __syncthreads();
  if (tid < R) {
    const double x = h_fin * wn_r;   // activations for the un-normalised basis
    h_s[tid] = x;
    hp_s[(tid & 7) * HF_HP + (tid >> 3)] = x;
  }
  __syncthreads();
  double* lam_d = recv;              // 4 partial groups of D_hat (the exchange buffers are idle after the last wait)
  {
    double x0 = 0.0, x1 = 0.0, y0 = 0.0, y1 = 0.0, d0 = 0.0, d1 = 0.0, e0 = 0.0, e1 = 0.0;
    const double* base = Ws + rq * 32 + ((2 * l16) ^ (kg << 1));
    const double* hp = hp_s + kg * HF_HP;
    int k = kg, j = 0;
#pragma unroll 4
    for (; k + 8 < R; k += 16, j += 2) {
      const double2 hh = *reinterpret_cast<const double2*>(hp + j);
      const double2 w0 = *reinterpret_cast<const double2*>(base + (size_t)k * HF_ROWS);
      const double2 w1 = *reinterpret_cast<const double2*>(base + (size_t)(k + 8) * HF_ROWS);
      const double hx0 = k < R1 ? hh.x : 0.0, hd0 = k < R1 ? 0.0 : hh.x;
      const double hx1 = k + 8 < R1 ? hh.y : 0.0, hd1 = k + 8 < R1 ? 0.0 : hh.y;
      x0 = fma(w0.x, hx0, x0);
      x1 = fma(w0.y, hx0, x1);
      d0 = fma(w0.x, hd0, d0);
      d1 = fma(w0.y, hd0, d1);
      y0 = fma(w1.x, hx1, y0);
      y1 = fma(w1.y, hx1, y1);
      e0 = fma(w1.x, hd1, e0);
      e1 = fma(w1.y, hd1, e1);
    }
    if (k < R) {
      const double h0 = hp[j];
      const double2 w0 = *reinterpret_cast<const double2*>(base + (size_t)k * HF_ROWS);
      const double hx0 = k < R1 ? h0 : 0.0, hd0 = k < R1 ? 0.0 : h0;
      x0 = fma(w0.x, hx0, x0);
      x1 = fma(w0.y, hx0, x1);
      d0 = fma(w0.x, hd0, d0);
      d1 = fma(w0.y, hd0, d1);
    }
    x0 += y0; x1 += y1; d0 += e0; d1 += e1;
    x0 += __shfl_xor_sync(0xffffffffu, x0, 16);
    x1 += __shfl_xor_sync(0xffffffffu, x1, 16);
    d0 += __shfl_xor_sync(0xffffffffu, d0, 16);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 16);
    if (half == 0) {
      *reinterpret_cast<double2*>(lam_part + kp * HF_ROWS + rq * 32 + 2 * l16) = make_double2(x0, x1);
      *reinterpret_cast<double2*>(lam_d + kp * HF_ROWS + rq * 32 + 2 * l16) = make_double2(d0, d1);
    }
    if (tail_rank && warp < E) {  // tail row `warp`: lanes over atoms
      double sx = 0.0, sd = 0.0;
      for (int kk = lane; kk < R; kk += 32) {
        const double t = Wt[(size_t)warp * R + kk], hv = h_s[kk];
        sx = fma(t, kk < R1 ? hv : 0.0, sx);
        sd = fma(t, kk < R1 ? 0.0 : hv, sd);
      }
      sx = warp_sum(sx);
      sd = warp_sum(sd);
      if (lane == 0) {
        misc[16 + warp] = sx;
        misc[24 + warp] = sd;
      }
    }
  }
  __syncthreads();
  {
    double* dx = st.Xhat + (size_t)slot * LDF;
    double* dd = st.Dhat + (size_t)slot * LDF;
    if (tid < HF_ROWS) {
      double sx = 0.0, sd = 0.0;
#pragma unroll
      for (int q = 0; q < HF_KG / 2; ++q) {
        sx += lam_part[q * HF_ROWS + tid];
        sd += lam_d[q * HF_ROWS + tid];
      }
      dx[f0 + tid] = sx;
      dd[f0 + tid] = sd;
    } else if (tail_rank && tid < HF_ROWS + E) {
      dx[HF_CL * HF_ROWS + tid - HF_ROWS] = misc[16 + tid - HF_ROWS];
      dd[HF_CL * HF_ROWS + tid - HF_ROWS] = misc[24 + tid - HF_ROWS];
    }
  }
  cluster.sync();  // nobody may exit while a peer can still read its exchange buffers
}

bool hsolve_fast_supported(snmfnat_ctx* ctx, const OnlineDims& d) {
  const int E = d.F - HF_CL * HF_ROWS;
  if (E < 0 || E > 8) return false;
  if (d.R > HF_THREADS || (d.R + 15) / 16 > HF_WARPS - 2) return false;  // two warps are kept for the cost / sum(h) roles
  return (int)hf_layout(d.F, d.R).bytes <= ctx->max_smem_optin;
}

void launch_hsolve_fast(snmfnat_ctx* ctx, const OnlineDims& d, const OnlineScalars& sc, const SlotState& st,
                        const FrameArrays& fr, const double* h_init, int n_active, int g_step) {
  const HfLayout L = hf_layout(d.F, d.R);
  SN_CUDA(cudaFuncSetAttribute(hsolve_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.bytes));
  hsolve_fast_kernel<<<dim3(HF_CL * n_active), dim3(HF_THREADS), L.bytes, ctx->stream>>>(d, sc, st, fr, h_init, g_step,
                                                                                         log_table(ctx));
  count_launch(ctx);
  check_launch(ctx, "hsolve_fast_kernel");
}

// =====================================================================================================
// W-solve, fast path
// =====================================================================================================
// Template parameters: KT = atom tiles of 8, CL = CTAs per cluster, TPC = full 8-row tiles (= warps) per CTA, VSMEM = the
// CTA's slice of V = lambda_d_blk is staged once in shared memory.  Shipped: <7 or 8, 4, 16, true>, one CTA per SM.  (An
// <., 8, 8, false> geometry -- V read from L2 once per pass, 82 KB of shared memory, two CTAs of different streams per SM --
// was measured 10 % slower in round 1: twice the ranks in every exchange.)

template <int KT, int CL, int TPC, bool VSMEM>
struct WfLayout {
  static constexpr int KMAX = KT * 8;
  static constexpr int XN = 3 * KMAX + 8;
  static constexpr int WARPS = TPC;          // one 8-row tensor-core tile per warp; the F % 8 leftover rows are shared out
  int NP, HSd, VS, VROWS;
  int LOWN;                                  // warps that own a 16-column group of the leftover rows
  size_t off_H, off_V, off_red, off_recv, off_hs, off_wn, off_tot, off_tab, off_scratch, off_bar, off_Wl, off_Gl, off_rl, off_Gp, bytes;
  __host__ __device__ WfLayout(int m_a, int nleft) {
    NP = (m_a + 15) / 16 * 16;
    LOWN = NP / 16 < WARPS ? NP / 16 : WARPS;
    HSd = NP + ((2 - NP % 8) + 8) % 8;          // == 2 (mod 8): conflict-free fragment loads in both GEMMs
    VROWS = (TPC + 1) * 8;                      // local rows (leftover tile included)
    VS = VROWS + 1;                             // odd stride: conflict-free (t, row) fragment reads (an even stride with
                                                // one bulk copy per history column was measured slower: 2-way conflicts)
    size_t o = 0;
    off_H = o;       o += (size_t)KMAX * HSd;
    off_V = o;       o += VSMEM ? (size_t)NP * VS : 0;
    off_red = o;     o += (size_t)2 * WARPS * KMAX;
    off_recv = o;    o += (size_t)2 * CL * XN;      // [2][CL][XN] partials pushed by the CTAs of the cluster
    off_hs = o;      o += KMAX;
    off_wn = o;      o += KMAX;
    off_tot = o;     o += 2 * KMAX;
    o = (o + 1) & ~(size_t)1;
    off_tab = o;     o += 256;
    off_scratch = o; o += 64;
    off_bar = o;     o += 2;                        // 2 mbarriers
    // leftover rows (F % 8) of W and G, and per owner warp the ratio of its 16 history columns and its partial of G
    off_Wl = o;      o += (size_t)nleft * KMAX;
    off_Gl = o;      o += (size_t)nleft * KMAX;
    off_rl = o;      o += (size_t)LOWN * nleft * 16;
    off_Gp = o;      o += (size_t)LOWN * nleft * KMAX;
    bytes = o * sizeof(double);
  }
};

// 8-byte asynchronous global -> shared copy; `valid == false` writes zero without reading
__device__ __forceinline__ void cp_async8(double* dst_smem, const double* src, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  const int sz = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ double rows8_sum_f(double v) {
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  return v;
}

// Phase probe (build with -DSNMFNAT_WS_PROBE): thread 0 of the first updating cluster's rank 0 accumulates clock64 deltas.
#ifdef SNMFNAT_WS_PROBE
__device__ unsigned long long g_ws_probe[16];
#define WS_TICK(i)                                                       \
  do {                                                                   \
    if (probe) {                                                         \
      const long long t_ = clock64();                                    \
      atomicAdd(&g_ws_probe[i], (unsigned long long)(t_ - tprev));       \
      tprev = t_;                                                        \
    }                                                                    \
  } while (0)
#else
#define WS_TICK(i) do {} while (0)
#endif

template <int KT, int CL, int TPC, bool VSMEM>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(WfLayout<KT, CL, TPC, VSMEM>::WARPS * 32, VSMEM ? 1 : 2)
wsolve_fast_kernel(OnlineDims d, OnlineScalars sc, SlotState st, TraceArrays tr, int has_trace, int g_step,
                   const double2* __restrict__ log_tab) {
  using LT = WfLayout<KT, CL, TPC, VSMEM>;
  constexpr int KMAX = LT::KMAX;
  constexpr int XN = LT::XN;
  constexpr int WARPS = LT::WARPS;
  constexpr int THREADS = WARPS * 32;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  // launch order written by this hop's gain kernel (gated slots first, longest expected solve first), if any
  int idx = (int)(blockIdx.x / CL);
  {
    const int pg = d.slot0 & 15;
    if (st.ws_perm && d.slot0 < 16 && st.ws_perm_step[pg] == g_step) idx = st.ws_perm[(size_t)pg * st.ms_perm_stride + idx];
  }
  const int slot = d.slot0 + idx * d.slot_stride;
  const int l = g_step + 1 - st.l_offset[slot];
  if (l < 1 || l > st.n_hops[slot]) return;
  if (!st.do_update[slot]) return;  // uniform over the cluster

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifdef SNMFNAT_WS_PROBE
#ifndef SNMFNAT_WS_PROBE_RANK
#define SNMFNAT_WS_PROBE_RANK 0
#endif
#ifndef SNMFNAT_WS_PROBE_TID
#define SNMFNAT_WS_PROBE_TID 0
#endif
  const bool probe = (blockIdx.x % (4 * 37)) == SNMFNAT_WS_PROBE_RANK && tid == SNMFNAT_WS_PROBE_TID;
  long long tprev = clock64();
#endif
  const int g = lane >> 2, tg = lane & 3;
  const int F = d.F, LDF = d.LDF, R_a = d.R_a, R_d = d.R_d, n = d.m_a;
  const double flr = sc.flr;
  const LT L(n, F % 8);
  const int HSd = L.HSd, NP = L.NP, VS = L.VS;

  extern __shared__ __align__(1024) double smem[];
  double* Hs = smem + L.off_H;
  double* Vs = smem + L.off_V;        // [NP][VS]: V slice of this CTA, pad = floor (VSMEM only)
  double* red = smem + L.off_red;     // [2][WARPS][KMAX]
  double* recv = smem + L.off_recv;   // [2][CL][XN]
  const unsigned bar0 = (unsigned)__cvta_generic_to_shared(smem + L.off_bar);
  if (tid == 0) {
    hf_mbar_init(bar0, 1);
    hf_mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  double* hs_s = smem + L.off_hs;
  double* wn_s = smem + L.off_wn;
  double* tot = smem + L.off_tot;     // [2][KMAX]
  double2* tab = reinterpret_cast<double2*>(smem + L.off_tab);
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

  // tiles: rank r owns full tiles [TPC r, TPC r + TPC), one per warp.  The leftover rows (F % 8; the Nyquist bin for
  // F = 513) belong to the last rank and are plain FMAs: the warp that owns history-column group tgp (the highest warp
  // indices: they leave the tile loop first) also works the leftover rows for those 16 columns inside its tile loop, and
  // the column sums join the CTA partial inside cluster_combine -- no extra barrier, nothing on the critical path.  (A
  // 17th warp for them costs every thread a quarter of its registers: a 17-warp CTA is allocated like 20 warps, 96
  // registers per thread instead of 128.)
  const int NFT = F / 8;
  const int nleft = F - NFT * 8;
  const int row0 = rank * TPC * 8;                     // first global row of this CTA's slice
  const int tile_local = warp;                         // local tile index 0..TPC-1
  const bool tile_valid = (rank * TPC + warp) < NFT;
  const bool has_left = (rank == CL - 1) && nleft > 0; // uniform over the CTA
  const int frow = row0 + tile_local * 8 + g;          // global row of this lane's fragment row
  const bool row_valid = tile_valid && frow < F;
  double* Wl = smem + L.off_Wl;                        // [nleft][KMAX]
  double* Gl = smem + L.off_Gl;                        // [nleft][KMAX]
  double* rlp = smem + L.off_rl;                       // [LOWN][nleft][16] ratio of the leftover rows, per owner warp
  double* Glp = smem + L.off_Gp;                       // [LOWN][nleft][KMAX] partial G of the leftover rows, per owner warp
  const int lown = L.LOWN;
  // local row index inside Vs: tiles 0..TPC-1 -> rows 0..8 TPC-1, leftover tile -> the 8 rows after them
  const int vrow = tile_local * 8 + g;

  // ---- staging.  Everything that comes from HBM is requested up front so that the latencies overlap: the W
  //      fragments (registers; two dependent loads), then the V slice and the raw history activations as asynchronous
  //      copies straight into shared memory (padding = 0; v = max(v, flr) is applied where V is read, sparse_nmf.m:169;
  //      H is scaled by the column norms in place once they are known) ----
  double w[KT][2], gacc[KT][2];
#pragma unroll
  for (int j = 0; j < KT; ++j)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int k = 8 * j + 2 * tg + e;
      double x = 0.0;
      if (row_valid && k < Ru) x = Bcur[(size_t)idx_up[k] * LDF + frow];
      w[j][e] = x;
      gacc[j][e] = 0.0;
    }
  if (has_left)
    for (int i = tid; i < nleft * KMAX; i += THREADS) {
      const int r = i / KMAX, k = i - r * KMAX;
      Wl[i] = (k < Ru) ? Bcur[(size_t)idx_up[k] * LDF + NFT * 8 + r] : 0.0;
    }
  if (tid < 128) tab[tid] = log_tab[tid];
  if (VSMEM) {
    for (int t = warp; t < NP; t += WARPS)
      for (int r = lane; r < L.VROWS; r += 32) {
        int fr_ = row0 + r;
        bool ok;
        if (r < TPC * 8) ok = fr_ < NFT * 8;
        else {
          fr_ = NFT * 8 + (r - TPC * 8);
          ok = (rank == CL - 1) && fr_ < F;
        }
        ok = ok && t < n;
        cp_async8(Vs + (size_t)t * VS + r, Vg + (ok ? (size_t)t * LDF + fr_ : 0), ok);
      }
  }
  for (int k = warp; k < KMAX; k += WARPS) {
    const bool kv = k < Ru;
    const int src = kv ? idx_up[k] : 0;
    for (int t = lane; t < NP; t += 32) {
      const bool ok = kv && t < n;
      cp_async8(Hs + (size_t)k * HSd + t, Adb + (ok ? (size_t)t * R_a + src : 0), ok);
    }
  }

  // per-warp partial of a per-column quantity -> red[which][warp][k]
  auto warp_partial = [&](int which, auto&& f) {
    // Sums over the 8 fragment rows (the lanes that share tg) of 2 KT per-lane values.  Instead of three shuffles per
    // value, eight values are reduced together: at every level a lane keeps the half of the values its own row bit selects
    // and hands the other half to its partner, so 4 + 2 + 1 exchanges finish eight sums, each ending up on one of the
    // eight lanes (same pairing order (g^1, g^2, g^4) as a per-value butterfly: bit-identical sums).
    double* rw = red + ((size_t)which * WARPS + warp) * KMAX;
    const bool b0 = g & 1, b1 = g & 2, b2 = g & 4;
#pragma unroll
    for (int base = 0; base < 2 * KT; base += 8) {
      double x[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = (base + i < 2 * KT) ? f((base + i) >> 1, (base + i) & 1) : 0.0;
      double y[4], z[2];
#pragma unroll
      for (int i = 0; i < 4; ++i) y[i] = (b0 ? x[i + 4] : x[i]) + __shfl_xor_sync(0xffffffffu, b0 ? x[i] : x[i + 4], 4);
#pragma unroll
      for (int i = 0; i < 2; ++i) z[i] = (b1 ? y[i + 2] : y[i]) + __shfl_xor_sync(0xffffffffu, b1 ? y[i] : y[i + 2], 8);
      const double s = (b2 ? z[1] : z[0]) + __shfl_xor_sync(0xffffffffu, b2 ? z[0] : z[1], 16);
      const int v = base + (b0 ? 4 : 0) + (b1 ? 2 : 0) + (b2 ? 1 : 0);     // the value whose sum this lane holds
      if (v < 2 * KT) rw[8 * (v >> 1) + 2 * tg + (v & 1)] = s;
    }
  };
  // what the leftover rows add to the CTA partial of column k (last rank only; evaluated inside cluster_combine by the
  // thread that sums column k): squares for the norms, or cw_k = sum w / s_k = sum G.*w, with G gathered from the owner
  // warps' partials and kept in Gl for the update
  enum { LEFT_SQ = 0, LEFT_SUMS = 1 };
  auto left_term = [&](int mode, int which, int k) {
    double s = 0.0;
    for (int r = 0; r < nleft; ++r) {
      const double wv = Wl[r * KMAX + k];
      if (mode == LEFT_SQ) s = fma(wv, wv, s);
      else if (which == 0) s += wv;
      else {
        double gsum = 0.0;
        for (int o = 0; o < lown; ++o) gsum += Glp[((size_t)o * nleft + r) * KMAX + k];
        Gl[r * KMAX + k] = gsum;
        s = fma(gsum, wv, s);
      }
    }
    return s;
  };
  // CTA partial (fixed warp order) PUSHED to the CTAs of the cluster (st.async, bytes counted on the receiver's
  // mbarrier: no cluster barrier / fence); totals in rank order -> tot[which][k]
  unsigned rnd = 0;
  auto cluster_combine = [&](int nwhich, bool with_cost, double* extra_out, int left_mode, bool inv_sqrt = false) {
    __syncthreads();
    const unsigned buf = rnd & 1u, parity = (rnd >> 1) & 1u;
    const unsigned bar = bar0 + 8u * buf;
    double* rb = recv + (size_t)buf * CL * XN;
    const int npay = nwhich * KMAX;
    const unsigned my_row = (unsigned)__cvta_generic_to_shared(rb + (size_t)rank * XN);
    if (tid == 0) hf_mbar_expect_tx(bar, (unsigned)(CL * (npay + 1) * sizeof(double)));
    if (tid < npay) {
      const int which = tid / KMAX, k = tid - which * KMAX;
      double s = 0.0;
#pragma unroll
      for (int ww = 0; ww < WARPS; ++ww) s += red[((size_t)which * WARPS + ww) * KMAX + k];
      if (has_left) s += left_term(left_mode, which, k);
      const unsigned la = my_row + 8u * (unsigned)tid;
#pragma unroll
      for (int c = 0; c < CL; ++c) hf_st_async(hf_mapa(la, c), s, hf_mapa(bar, c));
    } else if (tid == THREADS - 1) {  // per-warp cost partials, fixed order
      double s = 0.0;
      if (with_cost) {
#pragma unroll
        for (int ww = 0; ww < WARPS; ++ww) s += scratch[ww];
      }
      const unsigned la = my_row + 8u * (unsigned)(3 * KMAX);
#pragma unroll
      for (int c = 0; c < CL; ++c) hf_st_async(hf_mapa(la, c), s, hf_mapa(bar, c));
    }
    hf_mbar_wait(bar, parity);
    if (tid < npay) {
      double s = 0.0;
#pragma unroll
      for (int c = 0; c < CL; ++c) s += rb[(size_t)c * XN + tid];
      tot[tid] = inv_sqrt ? 1.0 / sqrt(s) : s;   // one sqrt + division per column instead of one per element
      if (inv_sqrt && has_left && tid < Ru)      // the leftover rows are normalised here, by the thread that owns atom k,
        for (int r = 0; r < nleft; ++r) Wl[r * KMAX + tid] *= tot[tid];   // so that the barrier below publishes them
    }
    if (extra_out) {
      double s = 0.0;
#pragma unroll
      for (int c = 0; c < CL; ++c) s += rb[(size_t)c * XN + 3 * KMAX];
      *extra_out = s;
    }
    ++rnd;
    __syncthreads();
  };
  WS_TICK(0);
  cluster.sync();  // the mbarriers of every CTA are initialised before anybody pushes
  WS_TICK(1);

  // column norms of init_w (sparse_nmf.m:158); the cluster barrier above also published Wl inside the CTA
  warp_partial(0, [&](int j, int e) { return w[j][e] * w[j][e]; });
  cluster_combine(1, false, nullptr, LEFT_SQ);
  if (tid < KMAX) wn_s[tid] = (tid < Ru) ? sqrt(tot[tid]) : 1.0;
  __syncthreads();
  if (has_left && tid < KMAX)
    for (int r = 0; r < nleft; ++r) Wl[r * KMAX + tid] = Wl[r * KMAX + tid] / wn_s[tid];
#pragma unroll
  for (int j = 0; j < KT; ++j)
#pragma unroll
    for (int e = 0; e < 2; ++e) w[j][e] = w[j][e] / wn_s[8 * j + 2 * tg + e];   // :159
  // H = init_h .* wn (:160), zero padded (each thread scales what it copied itself); row sums (constant over the solve)
  cp_async_wait_all();
  for (int k = warp; k < KMAX; k += WARPS) {
    const double wn = wn_s[k];
    for (int t = lane; t < NP; t += 32) Hs[(size_t)k * HSd + t] *= wn;
  }
  __syncthreads();   // also publishes every thread's part of the V slice
  if (tid < KMAX) {
    double s = 0.0;
    for (int t = 0; t < n; ++t) s += Hs[(size_t)tid * HSd + t];
    hs_s[tid] = s;
  }
  __syncthreads();
  double hsum_all = 0.0;
  for (int k = 0; k < Ru; ++k) hsum_all += hs_s[k];
  WS_TICK(2);

  // ---- multiplicative updates ----
  int it = 0;
  double last_cost = INFINITY, cost = 0.0;
  const int ngroups = NP / 16;
  const double* vbase = Vs + vrow;
  // V straight from L2 (!VSMEM): this lane's 4 history columns of group tgp, floored; padding = floor
  const double* vglob = Vg + (row_valid ? frow : 0);
  auto load_v4 = [&](int tgp, double (&dst)[4]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int t = tgp * 16 + 4 * tg + q;
      double x = flr;
      if (row_valid && t < n) x = fmax(__ldg(vglob + (size_t)t * LDF), flr);        // sparse_nmf.m:169
      dst[q] = x;
    }
  };
  double cacc = 0.0;
  bool want_cost = false;
  // One pass over the history columns, compiled once per number of atom tiles in use (kt = ceil(Ru / 8)): with the tile
  // count a compile-time constant the loop body is straight-line code and the fragment loads can move ahead of the mma.
  auto tile_pass = [&](auto ktc) {
    constexpr int KTC = decltype(ktc)::value;

    // a warp without a tile only takes part in the reductions (and in the leftover rows on the last rank)
    for (int tgp = 0; tgp < ((tile_valid || has_left) ? ngroups : 0); ++tgp) {
      const int n0 = tgp * 16;
      if (tile_valid) {
      double vcur[4];
      if (!VSMEM) load_v4(tgp, vcur);                                // in flight during GEMM 1 (and the other CTA's work)
      // GEMM 1: lambda tile = W * H for 16 history columns (even columns -> c0, odd -> c1)
      double c0[2] = {0.0, 0.0}, c1[2] = {0.0, 0.0};
#pragma unroll
      for (int j = 0; j < KTC; ++j) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const double2 b = *reinterpret_cast<const double2*>(Hs + (size_t)(8 * j + 2 * tg + e) * HSd + n0 + 2 * g);
            dmma884(c0[0], c0[1], w[j][e], b.x);
            dmma884(c1[0], c1[1], w[j][e], b.y);
          }
        }
      // ratio v./lambda and cost terms; this lane holds history columns n0+4tg+q, q=0..3
      double rt[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const double lam = fmax((q & 1) ? c1[q >> 1] : c0[q >> 1], flr);                 // :243
        const double v = VSMEM ? fmax(vbase[(size_t)(n0 + 4 * tg + q) * VS], flr) : vcur[q];
        const double r = v * fast_rcp(lam);
        if (want_cost) cacc += fma(v, fast_log(r, tab), lam - v);                         // :250
        rt[q] = r;
      }
      // GEMM 2: G tile += (v./lambda) * H'
#pragma unroll
      for (int j = 0; j < KTC; j += 2) {
          const double* hp = Hs + (size_t)(8 * j + g) * HSd + n0 + 4 * tg;
          const double2 a01 = *reinterpret_cast<const double2*>(hp);
          const double2 a23 = *reinterpret_cast<const double2*>(hp + 2);
          if (j + 1 < KTC) {
            const double2 b01 = *reinterpret_cast<const double2*>(hp + 8 * HSd);
            const double2 b23 = *reinterpret_cast<const double2*>(hp + 8 * HSd + 2);
            dmma884(gacc[j][0], gacc[j][1], rt[0], a01.x);
            dmma884(gacc[j + 1 < KTC ? j + 1 : j][0], gacc[j + 1 < KTC ? j + 1 : j][1], rt[0], b01.x);
            dmma884(gacc[j][0], gacc[j][1], rt[1], a01.y);
            dmma884(gacc[j + 1 < KTC ? j + 1 : j][0], gacc[j + 1 < KTC ? j + 1 : j][1], rt[1], b01.y);
            dmma884(gacc[j][0], gacc[j][1], rt[2], a23.x);
            dmma884(gacc[j + 1 < KTC ? j + 1 : j][0], gacc[j + 1 < KTC ? j + 1 : j][1], rt[2], b23.x);
            dmma884(gacc[j][0], gacc[j][1], rt[3], a23.y);
            dmma884(gacc[j + 1 < KTC ? j + 1 : j][0], gacc[j + 1 < KTC ? j + 1 : j][1], rt[3], b23.y);
          } else {
            dmma884(gacc[j][0], gacc[j][1], rt[0], a01.x);
            dmma884(gacc[j][0], gacc[j][1], rt[1], a01.y);
            dmma884(gacc[j][0], gacc[j][1], rt[2], a23.x);
            dmma884(gacc[j][0], gacc[j][1], rt[3], a23.y);
          }
        }
      }
      // leftover rows for the 16 history columns of this group, on the warp that owns the group
      if (has_left && (WARPS - 1 - tgp % WARPS) == warp) {
        const int o = tgp % WARPS;
        double* rlw = rlp + (size_t)o * nleft * 16;
        double* glw = Glp + (size_t)o * nleft * KMAX;
        const int t16 = lane & 15, kh = lane >> 4;
        const double* hcol = Hs + n0 + t16;
        for (int r = 0; r < nleft; ++r) {             // lambda and ratio: (column, half of the atoms) per lane
          const double* wr = Wl + r * KMAX;
          double a0 = 0.0, a1 = 0.0;
          int k = kh;
          for (; k + 2 < Ru; k += 4) {
            a0 = fma(wr[k], hcol[(size_t)k * HSd], a0);
            a1 = fma(wr[k + 2], hcol[(size_t)(k + 2) * HSd], a1);
          }
          if (k < Ru) a0 = fma(wr[k], hcol[(size_t)k * HSd], a0);
          double a = a0 + a1;
          a += __shfl_xor_sync(0xffffffffu, a, 16);
          if (kh == 0) {
            const int t = n0 + t16;
            double rr = 0.0;
            if (t < n) {
              const double lam = fmax(a, flr);
              const double v = fmax(Vs[(size_t)t * VS + TPC * 8 + r], flr);
              rr = v * fast_rcp(lam);
              if (want_cost) cacc += fma(v, fast_log(rr, tab), lam - v);
            }
            rlw[r * 16 + t16] = rr;
          }
        }
        __syncwarp();
        for (int k = lane; k < KMAX; k += 32) {       // this group's part of G: atom per lane (H is zero beyond Ru)
          const double* hk = Hs + (size_t)k * HSd + n0;
          for (int r = 0; r < nleft; ++r) {
            const double* rr = rlw + r * 16;
            double g0 = 0.0, g1 = 0.0;
#pragma unroll
            for (int q = 0; q < 16; q += 2) {
              g0 = fma(rr[q], hk[q], g0);
              g1 = fma(rr[q + 1], hk[q + 1], g1);
            }
            const double gp = g0 + g1;
            glw[r * KMAX + k] = (tgp < WARPS) ? gp : glw[r * KMAX + k] + gp;
          }
        }
        __syncwarp();
      }
    }
  };
  auto dispatch = [&](auto self, auto c) -> void {
    constexpr int C = decltype(c)::value;
    if (kt == C) tile_pass(c);
    else if constexpr (C > 0) self(self, std::integral_constant<int, C - 1>{});
  };
  for (;;) {
    cacc = 0.0;
    want_cost = sc.cost_check && it >= 1;
    dispatch(dispatch, std::integral_constant<int, KT>{});
    WS_TICK(3);
    // column reductions: cw_k = sum_f w, s_k = sum_f G.*w                               :215-221
    warp_partial(0, [&](int j, int e) { return w[j][e]; });
    warp_partial(1, [&](int j, int e) { return gacc[j][e] * w[j][e]; });
    cacc = warp_sum(cacc);
    if (lane == 0) scratch[warp] = cacc;
    double div = 0.0;
    cluster_combine(2, true, &div, LEFT_SUMS);
    WS_TICK(4);
#ifdef SNMFNAT_WS_PROBE
    if (probe) atomicAdd(&g_ws_probe[8], 1ull);
#endif
    bool stop = false;
    if (want_cost) {
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
    for (int j = 0; j < KT; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int k = 8 * j + 2 * tg + e;
        const double wv = w[j][e];
        const double hs = hs_s[k];
        const double dpw = fmax(hs + tot[KMAX + k] * wv, flr);
        const double dmw = gacc[j][e] + (hs * tot[k]) * wv;
        w[j][e] = (k < Ru) ? wv * dmw * fast_rcp(dpw) : 0.0;
        gacc[j][e] = 0.0;
      }
    if (has_left && tid < KMAX) {                  // thread k owns atom k of the leftover rows
      const int k = tid;
      const double hs = hs_s[k];
      for (int r = 0; r < nleft; ++r) {
        const double wv = Wl[r * KMAX + k];
        const double dpw = fmax(hs + tot[KMAX + k] * wv, flr);
        const double dmw = Gl[r * KMAX + k] + (hs * tot[k]) * wv;
        Wl[r * KMAX + k] = (k < Ru) ? wv * dmw * fast_rcp(dpw) : 0.0;
      }
    }
    WS_TICK(5);
    // column normalisation                                                              :242
    warp_partial(0, [&](int j, int e) { return w[j][e] * w[j][e]; });
    cluster_combine(1, false, nullptr, LEFT_SQ, true);
#pragma unroll
    for (int j = 0; j < KT; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int k = 8 * j + 2 * tg + e;
        if (k < Ru) w[j][e] = w[j][e] * tot[k];
      }
    ++it;
    WS_TICK(6);
  }
  WS_TICK(9);

  // ---- B_DFT_d = [B_rem, B_new, B_fix]  (bnmf_sep_event_RT_IS16.m:336) into the other buffer ----
  const int n_rem = R_a - Ru;
#pragma unroll
  for (int j = 0; j < KT; ++j)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int k = 8 * j + 2 * tg + e;
      if (row_valid && k < Ru) Bnext[(size_t)(n_rem + k) * LDF + frow] = w[j][e];
    }
  if (has_left)
    for (int i = tid; i < nleft * KMAX; i += THREADS) {
      const int r = i / KMAX, k = i - r * KMAX;
      if (k < Ru) Bnext[(size_t)(n_rem + k) * LDF + NFT * 8 + r] = Wl[i];
    }
  {
    // the not-updated adaptable atoms move to the front; columns >= R_a never change and are valid in both
    // buffers since the reset (state.cu), so they are not copied
    const int rb = F / CL, rr = F % CL;
    const int rows = rb + (rank < rr ? 1 : 0);
    const int r0 = rank * rb + (rank < rr ? rank : rr);
    for (int c = warp; c < n_rem; c += WARPS) {
      const double* __restrict__ src = Bcur + (size_t)idx_rem[c] * LDF + r0;
      double* __restrict__ dst = Bnext + (size_t)c * LDF + r0;
      for (int f = lane; f < rows; f += 32) dst[f] = src[f];
    }
  }
  if (rank == 0 && tid == 0) {
    st.w_iters[slot] = it;
    atomicAdd(&st.stats[2], (unsigned long long)it);
    if (has_trace) tr.info[(st.frame_base[slot] + g_step) * 4 + 3] = it;
  }
  WS_TICK(7);
  cluster.sync();  // peers may still be reading this CTA's exchange buffers
  WS_TICK(10);
#ifdef SNMFNAT_WS_PROBE
  if (probe) atomicAdd(&g_ws_probe[15], 1ull);
#endif
  if (rank == 0 && tid == 0) st.bd_sel[slot] = sel ^ 1;
}

static bool wsolve_geom_ok(const OnlineDims& d, int CL, int TPC) {
  return (d.F + 7) / 8 <= CL * TPC + 1 && d.F / 8 <= CL * TPC;
}

bool wsolve_fast_supported(snmfnat_ctx* ctx, const OnlineDims& d) {
  if (d.R_a > 64 || d.R_a < 1) return false;
  if (!wsolve_geom_ok(d, 4, 16)) return false;   // both geometries cover 64 full tiles + 1 leftover
  const size_t bytes = d.R_a <= 56 ? WfLayout<7, 4, 16, true>(d.m_a, d.F % 8).bytes : WfLayout<8, 4, 16, true>(d.m_a, d.F % 8).bytes;
  return (int)bytes <= ctx->max_smem_optin;
}

template <int KT, int CL, int TPC, bool VSMEM>
static void launch_wsolve_variant(snmfnat_ctx* ctx, const OnlineDims& d, const OnlineScalars& sc, const SlotState& st,
                                  const TraceArrays& t, int has_trace, int n_active, int g_step, const double2* tab) {
  const size_t smem = WfLayout<KT, CL, TPC, VSMEM>(d.m_a, d.F % 8).bytes;
  auto kern = wsolve_fast_kernel<KT, CL, TPC, VSMEM>;
  SN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (!VSMEM) SN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  kern<<<dim3(CL * n_active), dim3(WfLayout<KT, CL, TPC, VSMEM>::WARPS * 32), smem, ctx->stream>>>(d, sc, st, t, has_trace,
                                                                                                 g_step, tab);
}

void launch_wsolve_fast(snmfnat_ctx* ctx, const OnlineDims& d, const OnlineScalars& sc, const SlotState& st,
                        const TraceArrays* tr, int n_active, int g_step) {
  const double2* tab = log_table(ctx);
  TraceArrays t{};
  if (tr) t = *tr;
  if (d.R_a <= 56) launch_wsolve_variant<7, 4, 16, true>(ctx, d, sc, st, t, tr ? 1 : 0, n_active, g_step, tab);
  else launch_wsolve_variant<8, 4, 16, true>(ctx, d, sc, st, t, tr ? 1 : 0, n_active, g_step, tab);
  count_launch(ctx);
  check_launch(ctx, "wsolve_fast_kernel");
#ifdef SNMFNAT_WS_PROBE
  static int nl = 0;
  if (++nl == 140) {
    SN_CUDA(cudaStreamSynchronize(ctx->stream));
    unsigned long long pr[16];
    SN_CUDA(cudaMemcpyFromSymbol(pr, g_ws_probe, sizeof(pr)));
    const double nk = (double)(pr[15] ? pr[15] : 1), ni = (double)(pr[8] ? pr[8] : 1);
    fprintf(stderr, "snmfnat ws probe (%.0f solves, %.1f GEMM passes per solve), clk per solve: staging issue %.0f | cluster.sync %.0f | "
            "norms+H scaling %.0f | epilogue %.0f | final cluster.sync %.0f ; clk per pass: GEMMs %.0f | reductions+combine %.0f | "
            "W update %.0f | norm combine %.0f\n", nk, ni / nk, pr[0] / nk, pr[1] / nk, pr[2] / nk, (pr[9] + pr[7]) / nk, pr[10] / nk,
            pr[3] / ni, pr[4] / ni, pr[5] / ni, pr[6] / ni);
  }
#endif
}

}  // namespace snmfnat
