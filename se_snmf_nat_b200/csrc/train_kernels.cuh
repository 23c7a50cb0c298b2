// Device side of the frame-sharded dictionary training (sparse_nmf with W and H both updated, KL divergence,
// src/sparse_nmf.m:186-286 as called from run_basis_train.m:84-88): two warp-specialised tcgen05 kernels per
// multiplicative-update iteration (the small element-wise kernels live in train.cu).
//
//   hphase_kernel   one CTA owns a tile of 128 frames.  For every NC-bin chunk of the dictionary streamed through
//                   shared memory by TMA:   Lambda = H_tile * W_chunk'   (tcgen05.mma, tf32, accumulators in TMEM)
//                                           R      = V ./ Lambda         (epilogue warps, TMEM -> registers -> TMEM)
//                                           Num   += R * W_chunk         (tcgen05.mma, A operand = R read from TMEM)
//                   then H' = H .* Num ./ (colsum(W) + sparsity) written back through TMA.  The same pass yields the
//                   KL cost of the state it started from, sum(H',2) and the last (Nyquist) row of R*H''.
//   wphase_kernel   one CTA owns 128 bins of the dictionary (resident in shared memory) and streams NC-frame tiles
//                   of H':                  Lambda = W_rows * H'_tile'
//                                           R      = V ./ Lambda
//                                           G     += R * H'_tile         (accumulated in TMEM over all frames)
// Lambda and R never touch HBM.  The streamed tile is used K-major by the first product and MN-major by the second,
// which for 32-bit operands need different swizzles (umma.cuh), so TMA delivers it twice per stage.
#pragma once
#include <cstring>
#include "umma.cuh"

namespace snmfnat {
namespace train {

constexpr int BM = 128;      // rows of every MMA (TMEM lanes): frames (hphase) or bins (wphase)
constexpr int KB = 32;       // floats per swizzled column block (128 bytes)
constexpr int MAX_KP = 256;  // padded rank limit (TMEM: Kp accumulator columns + 2 Lambda buffers)
constexpr int TMEM_COLS = 512;
constexpr int LAM_COL = 256;
constexpr int THREADS = 32 * 10;  // warp 0: TMA, warp 1: MMA issue + TMEM alloc, warps 2-9: two epilogue groups
constexpr int MAX_STAGES = 4;
constexpr float FLRF = 1e-9f;

// H lives in HBM with half a tf32 ulp added to its bit pattern: the tensor core truncates fp32 operands to tf32, so
// it then sees round-to-nearest(h) instead of a value biased towards zero, and the exact fp32 state is recovered by
// subtracting the same constant (h >= 0 always; integer add / subtract is exactly reversible).
constexpr uint32_t H_BIAS = 0x1000u;
__host__ __device__ __forceinline__ float h_unbias(float x) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(__float_as_uint(x) - H_BIAS);
#else
  uint32_t u;
  memcpy(&u, &x, 4);
  u -= H_BIAS;
  memcpy(&x, &u, 4);
  return x;
#endif
}
__device__ __forceinline__ float h_bias(float x) { return __uint_as_float(__float_as_uint(x) + H_BIAS); }

struct HPhaseArgs {
  int F, Kp, nkb;      // bins, padded rank, Kp/32
  int nch, nlast;      // NC-bin chunks, MMA N of the last chunk (multiple of 16)
  int ntiles;          // ceil(T / 128)
  int nst;             // pipeline stages of the streamed dictionary chunk
  int update;          // 1: full H-update pass; 0: cost only
  int want_cost;
  int tail_row;        // F-1 when the last bin is handled outside the wphase MMA (F % 128 == 1), else -1
  long long T, ldt;
  const float* Vt;      // [F][ldt]  bin-major copy of V (coalesced for frame-per-thread reads)
  const float* invden;  // [Kp] 1 / max(colsum(W) + sparsity, flr)   (sparse_nmf.m:192-193)
  const float* wtail;   // [Kp] W(tail_row, :)
  float* hs_part;       // [grid][Kp] partial sum(H',2)
  float* gt_part;       // [grid][Kp] partial (V./Lambda')(tail_row,:) * H''
  double* cost_part;    // [grid]
  float* dbg;           // optional dump of CTA 0's first tile (diagnostics)
};

struct WPhaseArgs {
  int F, Kp, nkb;
  int nchunk, ngroups;  // 128-bin chunks; frame groups (grid = nchunk * ngroups)
  int nstages;          // ceil(T / NC)
  int nst;              // pipeline stages
  int ldv;
  long long T;
  const float* V;       // [T][ldv]  frame-major copy of V (coalesced for bin-per-thread reads)
  float* Gpart;         // [ngroups][nchunk*128][Kp]
  float* dbg;
};

__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ __forceinline__ uint8_t* align1024(uint8_t* p) {
  return (uint8_t*)(((uintptr_t)p + 1023) & ~(uintptr_t)1023);
}

template <int NC>
__host__ __device__ constexpr size_t hphase_smem_bytes(int nkb, int nst) {
  return (size_t)nkb * 16384 + (size_t)nst * 2 * nkb * NC * 128 + 2 * BM * 4 + 64 + 32 * 8 + 1024;
}
template <int NC>
__host__ __device__ constexpr size_t wphase_smem_bytes(int nkb, int nst) {
  return (size_t)nkb * 16384 + (size_t)nst * 2 * nkb * NC * 128 + 32 * 8 + 1024;
}

// ------------------------------------------------------------------------------------------------ H phase
template <int NC>
__global__ void __launch_bounds__(THREADS, 1)
hphase_kernel(const __grid_constant__ CUtensorMap mapH, const __grid_constant__ CUtensorMap mapHout,
              const __grid_constant__ CUtensorMap mapWk, const __grid_constant__ CUtensorMap mapWm, const HPhaseArgs a) {
  using namespace umma;
  constexpr int CBB = NC * 128;  // bytes of one column block of a streamed tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  const int nkb = a.nkb, Kp = a.Kp, nst = a.nst;
  const int tileb = nkb * CBB;                            // bytes of one copy of a streamed chunk
  uint8_t* Hs = smem;                                     // nkb x [128 x 128 B]
  uint8_t* Wk = Hs + nkb * 16384;                         // nst x nkb x [NC x 128 B]   SW128 (K-major use)
  uint8_t* Wm = Wk + nst * tileb;                         // nst x nkb x [NC x 128 B]   SW128_ATOM_32B (MN-major use)
  float* vtail_s = (float*)(Wm + nst * tileb);            // [128]
  float* dotp = vtail_s + BM;                             // [128]
  double* red = (double*)(dotp + BM);                     // [8]
  uint64_t* bars = (uint64_t*)(red + 8);
  uint64_t* h_full = bars + 0;
  uint64_t* h_empty = bars + 1;
  uint64_t* num_full = bars + 2;
  uint64_t* num_empty = bars + 3;
  uint64_t* lam_full = bars + 4;                 // [2]
  uint64_t* r_full = bars + 6;                   // [2]
  uint64_t* ws_full = bars + 8;                  // [MAX_STAGES]
  uint64_t* ws_empty = bars + 8 + MAX_STAGES;    // [MAX_STAGES]
  uint32_t* tmem_slot = (uint32_t*)(bars + 8 + 2 * MAX_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(h_full, 1);
    mbar_init(h_empty, 1);
    mbar_init(num_full, 1);
    mbar_init(num_empty, 2 * BM);
    for (int i = 0; i < 2; ++i) {
      mbar_init(lam_full + i, 1);
      mbar_init(r_full + i, BM);
    }
    for (int i = 0; i < MAX_STAGES; ++i) {
      mbar_init(ws_full + i, 1);
      mbar_init(ws_empty + i, 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int nch = a.nch;
  const bool upd = a.update != 0;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      tma_prefetch_desc(&mapH);
      tma_prefetch_desc(&mapWk);
      tma_prefetch_desc(&mapWm);
      uint32_t n = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++it) {
        const int t0 = tile * BM;
        bool h_loaded = false;
        for (int c = 0; c < nch; ++c, ++n) {
          if (c == nst) {  // the first dictionary chunks do not wait for the previous tile's write-back
            mbar_wait(h_empty, (it & 1) ^ 1);
            mbar_expect_tx(h_full, nkb * 16384);
            for (int kb = 0; kb < nkb; ++kb) tma_load_2d(Hs + kb * 16384, &mapH, h_full, kb * KB, t0);
            h_loaded = true;
          }
          const int s = n % nst;
          mbar_wait(ws_empty + s, ((n / nst) & 1) ^ 1);
          mbar_expect_tx(ws_full + s, (upd ? 2 : 1) * tileb);
          for (int kb = 0; kb < nkb; ++kb) tma_load_2d(Wk + s * tileb + kb * CBB, &mapWk, ws_full + s, kb * KB, c * NC);
          if (upd)
            for (int kb = 0; kb < nkb; ++kb) tma_load_2d(Wm + s * tileb + kb * CBB, &mapWm, ws_full + s, kb * KB, c * NC);
        }
        if (!h_loaded) {
          mbar_wait(h_empty, (it & 1) ^ 1);
          mbar_expect_tx(h_full, nkb * 16384);
          for (int kb = 0; kb < nkb; ++kb) tma_load_2d(Hs + kb * 16384, &mapH, h_full, kb * KB, t0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t id1 = idesc_tf32(BM, NC, 0, 0), id1l = idesc_tf32(BM, a.nlast, 0, 0), id2 = idesc_tf32(BM, Kp, 0, 1);
      const uint32_t hs_a = smem_u32(Hs), wk_a = smem_u32(Wk), wm_a = smem_u32(Wm);
      uint32_t n = 0;
      int it = 0;
      const bool probe = a.dbg && blockIdx.x == 0;
      long long p_h = 0, p_ws = 0, p_i1 = 0, p_r = 0, p_i2 = 0, p_t0 = clock64();
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++it) {
        long long q0 = clock64();
        mbar_wait(h_full, it & 1);
        p_h += clock64() - q0;
        tc_fence_after();
        const uint32_t nbase = n;
        for (int c = 0; c <= nch; ++c) {
          if (c < nch) {  // Lambda(c) = H_tile * W_chunk(c)'
            const uint32_t m = nbase + c, s = m % nst, b = m & 1;
            q0 = clock64();
            mbar_wait(ws_full + s, (m / nst) & 1);
            p_ws += clock64() - q0;
            q0 = clock64();
            tc_fence_after();
            const uint32_t d = tmem + LAM_COL + NC * b;
            const uint32_t id = (c == nch - 1) ? id1l : id1;
            for (int k = 0; k < Kp / 8; ++k) {
              const uint64_t da = smem_desc(hs_a + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
              const uint64_t db = smem_desc(wk_a + s * tileb + (k >> 2) * CBB + (k & 3) * 32, 16, 1024);
              mma_ss(d, da, db, id, k > 0);
            }
            mma_commit(lam_full + b);
            p_i1 += clock64() - q0;
          }
          if (c >= 1) {  // Num += R(c-1) * W_chunk(c-1)
            const uint32_t m = nbase + c - 1, s = m % nst, b = m & 1;
            q0 = clock64();
            mbar_wait(r_full + b, (m >> 1) & 1);
            p_r += clock64() - q0;
            q0 = clock64();
            tc_fence_after();
            if (upd) {
              if (c == 1) {
                mbar_wait(num_empty, (it & 1) ^ 1);
                tc_fence_after();
              }
              const int rows = (c - 1 == nch - 1) ? a.nlast : NC;
              for (int j = 0; j < rows / 8; ++j) {
                const uint64_t db = smem_desc(wm_a + s * tileb + j * 1024, CBB, 512, LAYOUT_SW128_32B);
                mma_ts(tmem, tmem + LAM_COL + NC * b + 8 * j, db, id2, (c > 1) || (j > 0));
              }
              mma_commit(ws_empty + s);
              if (c == nch) mma_commit(num_full);
            } else {
              mbar_arrive(ws_empty + s);
            }
            p_i2 += clock64() - q0;
          }
        }
        n = nbase + nch;
      }
      if (probe)
        printf("hphase probe (MMA issuer, CTA 0): total %lld clk, %d tiles x %d chunks, nst %d | wait h_full %lld, wait ws_full %lld, "
               "issue MMA1 %lld, wait r_full %lld, issue MMA2 %lld\n",
               clock64() - p_t0, it, nch, nst, p_h, p_ws, p_i1, p_r, p_i2);
    }
  } else {
    // ===================================================================== epilogue: two groups of 128 threads
    const int e = (warp - 2) >> 2;           // group: handles chunks with (global index & 1) == e
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int row = 32 * q + lane;           // frame inside the tile
    const int etid = (warp - 2) * 32 + lane; // 0..255
    const uint32_t lane_addr = tmem + ((uint32_t)(32 * q) << 16);
    float hs_acc = 0.f, gt_acc = 0.f;
    double cost_acc = 0.0;
    uint32_t n = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++it) {
      const long long t0 = (long long)tile * BM;
      const bool row_ok = (t0 + row) < a.T;
      const float* vcol = a.Vt + t0 + row;
      float cost_tile = 0.f;
      // V of this group's NEXT chunk is fetched while the current one is processed (the loads are the long pole of the
      // epilogue otherwise: one DRAM round trip per chunk with nothing else in flight)
      float vn[NC];
      auto load_v = [&](int c, float (&dst)[NC]) {
        const int f0 = c * NC;
#pragma unroll
        for (int j = 0; j < NC; ++j) dst[j] = (row_ok && f0 + j < a.F) ? __ldg(vcol + (size_t)(f0 + j) * a.ldt) : 0.f;
      };
      {
        const int c_first = ((int)(n & 1) == e) ? 0 : 1;
        if (c_first < nch) load_v(c_first, vn);
      }
      for (int c = 0; c < nch; ++c, ++n) {
        if ((int)(n & 1) != e) continue;
        const int f0 = c * NC;
        float v[NC];
#pragma unroll
        for (int j = 0; j < NC; ++j) v[j] = vn[j];
        if (c + 2 < nch) load_v(c + 2, vn);
        mbar_wait(lam_full + e, (n >> 1) & 1);
        tc_fence_after();
        uint32_t lam[NC];
        tmem_ld(lane_addr + LAM_COL + NC * e, lam);
        tmem_wait_ld();
        if (a.dbg && blockIdx.x == 0 && it == 0 && c == 0) {
          for (int j = 0; j < NC; ++j) {
            a.dbg[row * 32 + j] = __uint_as_float(lam[j]);
            a.dbg[4096 + row * 32 + j] = v[j];
          }
        }
        uint32_t rr[NC];
#pragma unroll
        for (int j = 0; j < NC; ++j) {
          const bool ok = row_ok && (f0 + j < a.F);
          const float vv = fmaxf(v[j], FLRF);                     // sparse_nmf.m:169
          const float ll = fmaxf(__uint_as_float(lam[j]), FLRF);  // :167,208
          const float r = __fdividef(vv, ll);
          rr[j] = ok ? to_tf32_rn(r) : 0u;
          if (a.want_cost && ok) cost_tile += vv * __logf(r) - vv + ll;  // :250
          v[j] = vv;
        }
        if (a.tail_row >= 0 && (a.tail_row / NC) == c) {
          float vt = 0.f;
#pragma unroll
          for (int j = 0; j < NC; ++j) vt = (j == (a.tail_row % NC)) ? v[j] : vt;
          vtail_s[row] = vt;
        }
        if (upd) {
          tmem_st(lane_addr + LAM_COL + NC * e, rr);
          tmem_wait_st();
        }
        tc_fence_before();
        mbar_arrive(r_full + e);
      }
      cost_acc += (double)cost_tile;
      if (upd) {
        // ---- H' = H .* Num ./ dph  (sparse_nmf.m:192-195), in place in the shared-memory tile
        mbar_wait(h_full, it & 1);
        mbar_wait(num_full, it & 1);
        tc_fence_after();
        float dot = 0.f;
        for (int kb = e; kb < nkb; kb += 2) {
          uint32_t num[32];
          tmem_ld32(lane_addr + kb * KB, num);
          uint8_t* hrow = Hs + kb * 16384 + row * 128;
          float4 hv[8];
#pragma unroll
          for (int g4 = 0; g4 < 8; ++g4) hv[g4] = *(const float4*)(hrow + (((g4 ^ row) & 7) << 4));
          tmem_wait_ld();
#pragma unroll
          for (int g4 = 0; g4 < 8; ++g4) {
            hv[g4].x = row_ok ? h_unbias(hv[g4].x) : 0.f;
            hv[g4].y = row_ok ? h_unbias(hv[g4].y) : 0.f;
            hv[g4].z = row_ok ? h_unbias(hv[g4].z) : 0.f;
            hv[g4].w = row_ok ? h_unbias(hv[g4].w) : 0.f;
          }
          if (a.dbg && blockIdx.x == 0 && it == 0) {
            for (int g4 = 0; g4 < 8; ++g4) {
              float* d0 = a.dbg + 8192 + row * Kp + kb * KB + 4 * g4;
              d0[0] = hv[g4].x; d0[1] = hv[g4].y; d0[2] = hv[g4].z; d0[3] = hv[g4].w;
              for (int u = 0; u < 4; ++u) d0[32768 + u] = __uint_as_float(num[4 * g4 + u]);
            }
          }
#pragma unroll
          for (int g4 = 0; g4 < 8; ++g4) {
            const int k = kb * KB + 4 * g4;
            const float4 id = __ldg((const float4*)(a.invden + k));
            const float4 wt = __ldg((const float4*)(a.wtail + k));
            float4 x = hv[g4];
            x.x = x.x * __uint_as_float(num[4 * g4 + 0]) * id.x;
            x.y = x.y * __uint_as_float(num[4 * g4 + 1]) * id.y;
            x.z = x.z * __uint_as_float(num[4 * g4 + 2]) * id.z;
            x.w = x.w * __uint_as_float(num[4 * g4 + 3]) * id.w;
            dot += wt.x * x.x + wt.y * x.y + wt.z * x.z + wt.w * x.w;
            x.x = h_bias(x.x); x.y = h_bias(x.y); x.z = h_bias(x.z); x.w = h_bias(x.w);
            *(float4*)(hrow + (((g4 ^ row) & 7) << 4)) = x;
          }
        }
        tc_fence_before();
        mbar_arrive(num_empty);
        fence_proxy_async();
        if (e == 1) dotp[row] = dot;
        named_bar_sync(1, 2 * BM);
        if (e == 0) {
          float rt = 0.f;
          if (a.tail_row >= 0 && row_ok) rt = __fdividef(vtail_s[row], fmaxf(dot + dotp[row], FLRF));
          dotp[row] = rt;
        }
        named_bar_sync(1, 2 * BM);
        if (etid == 0) {
          for (int kb = 0; kb < nkb; ++kb) tma_store_2d(&mapHout, Hs + kb * 16384, kb * KB, (int)t0);
          tma_store_commit();
        }
        // ---- column pass over the tile: sum(H',2) and the tail row of (V./Lambda') * H''
        if (etid < Kp) {
          const uint8_t* col = Hs + (etid >> 5) * 16384;
          const uint32_t j = etid & 31;
          float s1 = 0.f, s2 = 0.f;
          const int nrows = (int)((a.T - t0 < BM) ? (a.T - t0) : BM);
#pragma unroll 4
          for (int r = 0; r < nrows; ++r) {
            const float hval = h_unbias(*(const float*)(col + sw128_off(r, j)));
            s1 += hval;
            s2 += dotp[r] * hval;
          }
          hs_acc += s1;
          gt_acc += s2;
        }
        named_bar_sync(1, 2 * BM);
        if (etid == 0) {
          tma_store_wait_read();
          mbar_arrive(h_empty);
        }
      } else {
        named_bar_sync(1, 2 * BM);
        if (etid == 0) mbar_arrive(h_empty);
      }
    }
    if (etid == 0) tma_store_wait_all();
    if (etid < Kp) {
      a.hs_part[(size_t)blockIdx.x * Kp + etid] = hs_acc;
      a.gt_part[(size_t)blockIdx.x * Kp + etid] = gt_acc;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cost_acc += __shfl_xor_sync(0xffffffffu, cost_acc, o);
    if (lane == 0) red[warp - 2] = cost_acc;
    named_bar_sync(1, 2 * BM);
    if (etid == 0) {
      double t = 0.0;
      for (int i = 0; i < 8; ++i) t += red[i];
      a.cost_part[blockIdx.x] = t;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------ W phase
template <int NC>
__global__ void __launch_bounds__(THREADS, 1)
wphase_kernel(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapHk,
              const __grid_constant__ CUtensorMap mapHm, const WPhaseArgs a) {
  using namespace umma;
  constexpr int CBB = NC * 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  const int nkb = a.nkb, Kp = a.Kp, nst = a.nst;
  const int tileb = nkb * CBB;
  uint8_t* Wc = smem;                                  // nkb x [128 x 128 B]   resident dictionary rows
  uint8_t* Hk = Wc + nkb * 16384;                      // nst x nkb x [NC x 128 B]   SW128
  uint8_t* Hm = Hk + nst * tileb;                      // nst x nkb x [NC x 128 B]   SW128_ATOM_32B
  uint64_t* bars = (uint64_t*)(Hm + nst * tileb);
  uint64_t* wc_full = bars + 0;
  uint64_t* g_full = bars + 1;
  uint64_t* lam_full = bars + 2;                 // [2]
  uint64_t* r_full = bars + 4;                   // [2]
  uint64_t* hs_full = bars + 6;                  // [MAX_STAGES]
  uint64_t* hs_empty = bars + 6 + MAX_STAGES;    // [MAX_STAGES]
  uint32_t* tmem_slot = (uint32_t*)(bars + 6 + 2 * MAX_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(wc_full, 1);
    mbar_init(g_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(lam_full + i, 1);
      mbar_init(r_full + i, BM);
    }
    for (int i = 0; i < MAX_STAGES; ++i) {
      mbar_init(hs_full + i, 1);
      mbar_init(hs_empty + i, 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int chunk = blockIdx.x % a.nchunk, grp = blockIdx.x / a.nchunk;
  const int n_my = (a.nstages > grp) ? (a.nstages - grp + a.ngroups - 1) / a.ngroups : 0;

  if (warp == 0) {
    if (lane == 0 && n_my > 0) {
      tma_prefetch_desc(&mapW);
      tma_prefetch_desc(&mapHk);
      tma_prefetch_desc(&mapHm);
      mbar_expect_tx(wc_full, nkb * 16384);
      for (int kb = 0; kb < nkb; ++kb) tma_load_2d(Wc + kb * 16384, &mapW, wc_full, kb * KB, chunk * BM);
      for (int i = 0; i < n_my; ++i) {
        const int s = i % nst, st = grp + i * a.ngroups;
        mbar_wait(hs_empty + s, ((i / nst) & 1) ^ 1);
        mbar_expect_tx(hs_full + s, 2 * tileb);
        for (int kb = 0; kb < nkb; ++kb) tma_load_2d(Hk + s * tileb + kb * CBB, &mapHk, hs_full + s, kb * KB, st * NC);
        for (int kb = 0; kb < nkb; ++kb) tma_load_2d(Hm + s * tileb + kb * CBB, &mapHm, hs_full + s, kb * KB, st * NC);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && n_my > 0) {
      const uint32_t id3 = idesc_tf32(BM, NC, 0, 0), id4 = idesc_tf32(BM, Kp, 0, 1);
      const uint32_t wc_a = smem_u32(Wc), hk_a = smem_u32(Hk), hm_a = smem_u32(Hm);
      mbar_wait(wc_full, 0);
      tc_fence_after();
      for (int i = 0; i <= n_my; ++i) {
        if (i < n_my) {  // Lambda(i) = W_rows * H'_tile(i)'
          const int s = i % nst, b = i & 1;
          mbar_wait(hs_full + s, (i / nst) & 1);
          tc_fence_after();
          const uint32_t d = tmem + LAM_COL + NC * b;
          for (int k = 0; k < Kp / 8; ++k) {
            const uint64_t da = smem_desc(wc_a + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
            const uint64_t db = smem_desc(hk_a + s * tileb + (k >> 2) * CBB + (k & 3) * 32, 16, 1024);
            mma_ss(d, da, db, id3, k > 0);
          }
          mma_commit(lam_full + b);
        }
        if (i >= 1) {  // G += R(i-1) * H'_tile(i-1)
          const int m = i - 1, s = m % nst, b = m & 1;
          mbar_wait(r_full + b, (m >> 1) & 1);
          tc_fence_after();
          for (int j = 0; j < NC / 8; ++j) {
            const uint64_t db = smem_desc(hm_a + s * tileb + j * 1024, CBB, 512, LAYOUT_SW128_32B);
            mma_ts(tmem, tmem + LAM_COL + NC * b + 8 * j, db, id4, (m > 0) || (j > 0));
          }
          mma_commit(hs_empty + s);
          if (m == n_my - 1) mma_commit(g_full);
        }
      }
    }
  } else {
    const int e = (warp - 2) >> 2, q = warp & 3;
    const int row = 32 * q + lane;
    const int f = chunk * BM + row;
    const bool f_ok = f < a.F;
    const uint32_t lane_addr = tmem + ((uint32_t)(32 * q) << 16);
    float vn[NC];
    auto load_v = [&](int i, float (&dst)[NC]) {
      const long long t0 = (long long)(grp + i * a.ngroups) * NC;
#pragma unroll
      for (int j = 0; j < NC; ++j) dst[j] = (f_ok && t0 + j < a.T) ? __ldg(a.V + (size_t)(t0 + j) * a.ldv + f) : 0.f;
    };
    if (e < n_my) load_v(e, vn);
    for (int i = e; i < n_my; i += 2) {
      const long long t0 = (long long)(grp + i * a.ngroups) * NC;
      float v[NC];
#pragma unroll
      for (int j = 0; j < NC; ++j) v[j] = vn[j];
      if (i + 2 < n_my) load_v(i + 2, vn);   // next stage of this group, in flight while this one is processed
      mbar_wait(lam_full + e, (i >> 1) & 1);
      tc_fence_after();
      uint32_t lam[NC];
      tmem_ld(lane_addr + LAM_COL + NC * e, lam);
      tmem_wait_ld();
      if (a.dbg && blockIdx.x == 0 && i == 0) {
        for (int j = 0; j < NC; ++j) {
          a.dbg[row * 32 + j] = __uint_as_float(lam[j]);
          a.dbg[4096 + row * 32 + j] = v[j];
        }
      }
      uint32_t rr[NC];
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        const bool ok = f_ok && (t0 + j < a.T);
        const float r = __fdividef(fmaxf(v[j], FLRF), fmaxf(__uint_as_float(lam[j]), FLRF));
        rr[j] = ok ? to_tf32_rn(r) : 0u;
      }
      tmem_st(lane_addr + LAM_COL + NC * e, rr);
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(r_full + e);
    }
    // ---- G tile -> this group's partial in HBM
    float* gout = a.Gpart + ((size_t)grp * a.nchunk * BM + (size_t)chunk * BM + row) * Kp;
    if (n_my > 0) {
      mbar_wait(g_full, 0);
      tc_fence_after();
    }
    for (int kb = e; kb < nkb; kb += 2) {
      uint32_t g[32];
      if (n_my > 0) {
        tmem_ld32(lane_addr + kb * KB, g);
        tmem_wait_ld();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) g[j] = 0u;
      }
#pragma unroll
      for (int g4 = 0; g4 < 8; ++g4)
        *(uint4*)(gout + kb * KB + 4 * g4) = make_uint4(g[4 * g4], g[4 * g4 + 1], g[4 * g4 + 2], g[4 * g4 + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, TMEM_COLS);
  }
}

}  // namespace train
}  // namespace snmfnat
