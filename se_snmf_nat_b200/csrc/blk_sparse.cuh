// blk_sparse as a block-wide device function shared by the gain kernel and the stand-alone L1 entry.
#pragma once
#include "online.cuh"

namespace snmfnat {

// [Q, r_blk_out] = blk_sparse(X, D, r_blk, l, p)                                  src/blk_sparse.m:3-36
//   rb   : ring buffer [P_len_l][ld] of the normalised local SNR; slot `cur` receives this hop's column and the
//          chronological order (oldest first) is cur+1, cur+2, ..., cur (mod P_len_l)
//   Q_s  : out, [F] (shared memory); rs1, rs2, P_s: [F] scratch (shared memory); scratch: >= 32 doubles
// Every thread of the block must call it; it ends with the result visible to the whole block.
__device__ __forceinline__ void blk_sparse_dev(const double* __restrict__ Xh, const double* __restrict__ Dh, double* rb,
                                               int ld, int F, int l, int cur, const OnlineScalars& sc, double* Q_s,
                                               double* rs1, double* rs2, double* P_s, double* scratch) {
  const int tid = threadIdx.x, nt = blockDim.x, PL = sc.P_len_l;
  const double flr = sc.flr;
  double mx = -INFINITY;
  for (int f = tid; f < F; f += nt) {
    const double s = Xh[f] / fmax(Dh[f], flr);                     // :10
    rs1[f] = s;
    mx = fmax(mx, s);
  }
  mx = block_max(mx, scratch);
  for (int f = tid; f < F; f += nt) rb[(size_t)cur * ld + f] = rs1[f] / mx;   // :12,14
  for (int f = tid; f < F; f += nt) Q_s[f] = (f < sc.DCbin) ? 0.0 : 0.1;       // :16
  __syncthreads();
  if (l > PL) {                                                     // :19
    for (int f = tid; f < F; f += nt) {
      double a = 0.0, b = 0.0;
      for (int i = 1; i <= PL; ++i) {  // oldest column first
        const double x = rb[(size_t)((cur + i) % PL) * ld + f];
        a += x;
        b = fma(x, x, b);
      }
      rs1[f] = a;
      rs2[f] = b;
    }
    __syncthreads();
    const int k2 = sc.P_len_k / 2;
    const int kfirst = k2 + sc.DCbin;  // 1-based centre of the first window (:20)
    const int nwin = (F - k2 >= kfirst) ? (F - k2 - kfirst) / sc.blk_gap + 1 : 0;
    const double sqn = sqrt((double)(sc.P_len_l * sc.P_len_k));
    for (int w = tid; w < nwin; w += nt) {
      const int k = kfirst + w * sc.blk_gap;
      double l1 = 0.0, l2 = 0.0;
      for (int f = k - k2; f < k + k2; ++f) {  // rows k-k2+1 .. k+k2 (1-based)
        l1 += rs1[f];
        l2 += rs2[f];
      }
      P_s[w] = (sqn - l1 / sqrt(l2)) / (sqn - 1.0);                // :26 (Hoyer)
    }
    __syncthreads();
    if (tid == 0) {  // the fill is order dependent: Q(k-1) may have been written by the previous window (:28-30)
      const int g2 = (sc.blk_gap - 1) / 2;
      for (int w = 0; w < nwin; ++w) {
        const int k = kfirst + w * sc.blk_gap;
        const double pv = sc.alpha_p * Q_s[k - 2] + (1.0 - sc.alpha_p) * P_s[w];
        for (int q = k - 1 - g2; q <= k - 1 + g2; ++q)
          if (q >= 0 && q < F) Q_s[q] = pv;
      }
      const double q0 = Q_s[sc.P_len_k + sc.DCbin - 1];            // :32
      for (int q = 0; q < sc.P_len_k - 1 && q < F; ++q) Q_s[q] = q0;
    }
    __syncthreads();
  }
  for (int f = tid; f < sc.DCbin && f < F; f += nt) Q_s[f] = 0.0;    // :36
  __syncthreads();
}

}  // namespace snmfnat
