// Semi-supervised separation solve: p.basis_update_N / p.basis_update_E of src/bnmf_sep_event_RT_IS16.m:125-139
// (settings/bak_IS16_results/initial_setting_semisupervised.m:109).  The per-hop sparse_nmf call then runs with
// w_update_ind = [false(R_x,1); true(R_d,1)] (N) or [true(R_x,1); false(R_d,1)] (E): inside the solve the chosen half of
// the dictionary follows the frame (sparse_nmf.m:212-244, one column of V), the updated W is DISCARDED by the caller
// ([~, A] = sparse_nmf(...), :150-154) and the separated spectra are rebuilt from g.B_DFT_x / g.B_DFT_d with the A that
// came out of the joint iteration (:174,197).
//
// One CTA per stream.  The updated atoms need a private, normalised copy per stream: it lives in a per-slot scratch in
// global memory (SlotState::semi_w, F x (upd1-upd0) doubles, L2-resident while the solve runs), the other atoms are
// read un-normalised with the column scaling of sparse_nmf.m:157-160 folded into the small vectors as in
// hsolve_stream_kernel.  With one frame the W-update is a rank-one expression per atom:
//     rh(:,k) = (v./lambda) * h_k          dpw = max(h_k + (h_k * sum(r.*w_k)) * w_k, flr)
//     sum(h,2) = h_k                        dmw = r*h_k + (h_k * sum(w_k)) * w_k
// followed by the re-normalisation of :242.  Order of an iteration as in the reference: H-update, Lambda, W-update,
// normalise, Lambda, cost, stop rule (:186-285).  A completeness path for the reference's comparison settings, not a
// throughput path: the shipped settings never take it.
#include "online.cuh"

namespace snmfnat {

namespace {

constexpr int SE_THREADS = 576;   // >= F for F = 513: thread <-> row in the Lambda passes
constexpr int SE_WARPS = SE_THREADS / 32;

__global__ void __launch_bounds__(SE_THREADS)
hsolve_semi_kernel(OnlineDims d, OnlineScalars sc, SlotState st, FrameArrays fr, const double* __restrict__ h_init, int g_step) {
  const int slot = d.slot0 + (int)blockIdx.x * d.slot_stride;
  const int l = g_step + 1 - st.l_offset[slot];
  if (l < 1 || l > st.n_hops[slot]) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int F = d.F, R = d.R, R1 = d.R_x, LDF = d.LDF, u0 = d.upd0, u1 = d.upd1;
  const double flr = sc.flr;
  extern __shared__ __align__(16) double smem[];
  double* h_s = smem;              // [R] activations (normalised-basis convention, sparse_nmf.m:160)
  double* ht_s = h_s + R;          // [R] what multiplies the stored columns: h ./ wn (fixed atoms), h (private atoms)
  double* sc_s = ht_s + R;         // [R] 1 / wn of the stored column (1 for the private, already normalised copies)
  double* cs_s = sc_s + R;         // [R] column sum of the normalised atom (:192)
  double* g_s = cs_s + R;          // [R]
  double* v_s = g_s + R;           // [F]
  double* r_s = v_s + F;           // [F]
  double* scratch = r_s + F;       // [64]
  const double* __restrict__ W1 = st.Bx;
  const double* __restrict__ W2 = st.Bd[st.bd_sel[slot]] + (size_t)slot * d.R_d * LDF;
  double* Wp = st.semi_w + (size_t)slot * (u1 - u0) * LDF;
  auto raw = [&](int k) { return k < R1 ? W1 + (size_t)k * LDF : W2 + (size_t)(k - R1) * LDF; };
  auto col = [&](int k) -> const double* { return (k >= u0 && k < u1) ? Wp + (size_t)(k - u0) * LDF : raw(k); };
  const double* __restrict__ V = fr.Ym + (size_t)(st.frame_base[slot] + g_step) * LDF;

  // sparse_nmf.m:157-160: wn = sqrt(sum(w.^2)); w = w ./ wn; h = h .* wn'
  for (int k = warp; k < R; k += SE_WARPS) {
    const double* c = raw(k);
    const bool priv = k >= u0 && k < u1;
    double s1 = 0.0, s2 = 0.0;
    for (int f = lane; f < F; f += 32) {
      const double x = c[f];
      s1 += x;
      s2 = fma(x, x, s2);
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    const double wn = sqrt(s2);
    double csn = s1 / wn;
    if (priv) {
      double* o = Wp + (size_t)(k - u0) * LDF;
      double cs = 0.0;
      for (int f = lane; f < F; f += 32) {
        const double x = c[f] / wn;
        o[f] = x;
        cs += x;
      }
      csn = warp_sum(cs);
    }
    if (lane == 0) {
      sc_s[k] = priv ? 1.0 : 1.0 / wn;
      cs_s[k] = csn;
      h_s[k] = h_init[k] * wn;
    }
  }
  for (int f = tid; f < F; f += SE_THREADS) v_s[f] = fmax(V[f], flr);          // sparse_nmf.m:169
  __syncthreads();

  auto row_dot = [&](const double* x, int k0, int k1, int f, bool stored) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int k = k0;
    for (; k + 4 <= k1; k += 4) {
      a0 = fma((stored ? col(k) : raw(k))[f], x[k], a0);
      a1 = fma((stored ? col(k + 1) : raw(k + 1))[f], x[k + 1], a1);
      a2 = fma((stored ? col(k + 2) : raw(k + 2))[f], x[k + 2], a2);
      a3 = fma((stored ? col(k + 3) : raw(k + 3))[f], x[k + 3], a3);
    }
    for (; k < k1; ++k) a0 = fma((stored ? col(k) : raw(k))[f], x[k], a0);
    return (a0 + a1) + (a2 + a3);
  };
  // lambda = max(w*h, flr) (:167,196,243), r = v ./ lambda; returns this thread's share of the divergence (:250)
  auto lambda_pass = [&](bool want_div) {
    for (int k = tid; k < R; k += SE_THREADS) ht_s[k] = h_s[k] * sc_s[k];
    __syncthreads();
    double cterm = 0.0;
    for (int f = tid; f < F; f += SE_THREADS) {
      const double lam = fmax(row_dot(ht_s, 0, R, f, true), flr);
      const double v = v_s[f];
      r_s[f] = v / lam;
      if (want_div) cterm += v * log(v / lam) - v + lam;
    }
    __syncthreads();
    return cterm;
  };

  int it = 0;
  double last_cost = INFINITY, cost = 0.0;
  lambda_pass(false);
  while (it < sc.max_iter) {
    // ---- H-update (:189-196)
    for (int k = warp; k < R; k += SE_WARPS) {
      const double* c = col(k);
      double s = 0.0;
      for (int f = lane; f < F; f += 32) s = fma(c[f], r_s[f], s);
      s = warp_sum(s);
      if (lane == 0) g_s[k] = s * sc_s[k];
    }
    __syncthreads();
    for (int k = tid; k < R; k += SE_THREADS) h_s[k] = h_s[k] * g_s[k] / fmax(cs_s[k] + sc.sparsity, flr);
    __syncthreads();
    lambda_pass(false);
    // ---- W-update of the atoms [u0, u1) (:212-241, KL) and their re-normalisation (:242)
    for (int k = u0 + warp; k < u1; k += SE_WARPS) {
      double* c = Wp + (size_t)(k - u0) * LDF;
      const double hk = h_s[k];
      double a = 0.0, b = 0.0;
      for (int f = lane; f < F; f += 32) {
        const double w = c[f];
        a = fma(r_s[f], w, a);
        b += w;
      }
      const double s1 = hk * warp_sum(a), s2 = hk * warp_sum(b);
      double nn = 0.0;
      for (int f = lane; f < F; f += 32) {
        const double w = c[f];
        const double x = w * fma(s2, w, r_s[f] * hk) / fmax(fma(s1, w, hk), flr);
        c[f] = x;
        nn = fma(x, x, nn);
      }
      const double nrm = sqrt(warp_sum(nn));
      double cs = 0.0;
      for (int f = lane; f < F; f += 32) {
        const double x = c[f] / nrm;
        c[f] = x;
        cs += x;
      }
      cs = warp_sum(cs);
      if (lane == 0) cs_s[k] = cs;
    }
    __syncthreads();   // the rewritten columns are visible to the whole CTA
    ++it;
    // ---- cost of the new iterate and the stop rule (:248-283)
    const double cterm = lambda_pass(sc.cost_check != 0);
    if (sc.cost_check) {
      double hpart = 0.0;
      for (int k = tid; k < R; k += SE_THREADS) hpart += h_s[k];
      const double div = block_sum(cterm, scratch);
      const double hsum = block_sum(hpart, scratch);
      cost = div + sc.sparsity * hsum;                                           // :261
      bool stop = false;
      if (it > 1 && sc.conv_eps > 0.0) stop = fabs(cost - last_cost) / last_cost < sc.conv_eps;   // :274
      last_cost = cost;
      if (stop) break;
    }
  }
  __syncthreads();
  for (int k = tid; k < R; k += SE_THREADS) st.A[(size_t)slot * R + k] = h_s[k];
  if (tid == 0) {
    st.h_iters[slot] = it;
    st.h_cost[slot] = cost;
  }
  // reconstructions with the dictionaries of g, not with the W of the solve (bnmf_sep_event_RT_IS16.m:141-143,174,197)
  for (int f = tid; f < F; f += SE_THREADS) {
    st.Xhat[(size_t)slot * LDF + f] = row_dot(h_s, 0, R1, f, false);
    st.Dhat[(size_t)slot * LDF + f] = row_dot(h_s, R1, R, f, false);
  }
}

size_t semi_smem(const OnlineDims& d) { return ((size_t)5 * d.R + 2 * d.F + 64) * sizeof(double); }

}  // namespace

void launch_hsolve_semi(snmfnat_ctx* ctx, const OnlineDims& d, const OnlineScalars& sc, const SlotState& st,
                        const FrameArrays& fr, const double* h_init, int n_active, int g_step) {
  SN_REQUIRE(st.semi_w != nullptr && d.upd1 > d.upd0, SNMFNAT_EINVAL, "semi-supervised solve without its scratch");
  const size_t sm = semi_smem(d);
  SN_REQUIRE((int)sm <= ctx->max_smem_optin, SNMFNAT_EUNSUPPORTED,
             "semi-supervised H-solve: R_x+R_d = %d needs %zu bytes of shared memory per CTA", d.R, sm);
  if (sm > 48 * 1024) SN_CUDA(cudaFuncSetAttribute(hsolve_semi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  hsolve_semi_kernel<<<dim3(n_active), dim3(SE_THREADS), sm, ctx->stream>>>(d, sc, st, fr, h_init, g_step);
  count_launch(ctx);
  check_launch(ctx, "hsolve_semi_kernel");
}

}  // namespace snmfnat
