// The two warp-specialised tcgen05 kernels of a dictionary-training iteration (second generation; shared constants and
// helpers in train_kernels.cuh).
//
// What changed, and why (measured on B200, profiles/r01_train_*): with both operands in shared memory a tcgen05.mma
// of M = 128 reads its A operand at one 128-byte row per clock, i.e. ~128 clk per instruction whatever N is, while the
// tensor pipe itself needs N/2 clk.  The first-generation kernels streamed the other matrix in 16-row chunks, so the
// first product (Lambda = A * chunk') ran with N = 16: 97 clk per instruction measured, 6-9 % of the tensor pipe.
// Here:
//   * the first product is issued with N = 128: Lambda buffer [128 x 128] = A_tile [128 x Kp] * X_half' with X_half =
//     128 rows of the streamed matrix, delivered as K-major tiles [128 rows x 32 atoms] (16 KB),
//   * there are TWO Lambda buffers in tensor memory (2 x 128 columns beside the <= 256 accumulator columns), so the
//     first product of half c+1 runs while the epilogue warps turn half c into R = V ./ Lambda in place (V is
//     prefetched from HBM one 32-column chunk ahead) and the second product Acc += R * X_half of half c-1 follows,
//   * one ring of 16 KB units feeds both products in consumption order (a K-major tile or an MN-major 16-row slice per
//     unit), so all the shared memory beside the resident tile is in flight for whichever product is running.
// In the H phase the Nyquist bin (F % 128 == 1) stays off the tensor cores: its Lambda is a dot product per frame in
// the epilogue and its contribution to the numerator a rank-1 term of the H update.
#pragma once
#include "train_kernels.cuh"

namespace snmfnat {
namespace train {

constexpr int HB = 128;        // columns of one Lambda buffer (N of the first product)
constexpr int SL = 16;         // rows of one MN-major slice of the second product
constexpr int UNIT = 16384;    // ring unit: one K-major tile [128 x 128 B] or one slice [nkb][16 x 128 B]
constexpr int NU_MAX = 8;

struct HPhase2Args {
  int F, Fm, Kp, nkb;  // bins, bins that go through the tensor cores, padded rank, Kp/32
  int nh, nlast;       // 128-bin halves, N of the last one (multiple of 16)
  int ntiles;
  int update, want_cost;
  int tail_row;        // F-1 when that bin is handled by the epilogue (F % 128 == 1), else -1
  long long T, ldt;
  const float* Vt;      // [F][ldt]
  const float* invden;  // [Kp]
  const float* wtail;   // [Kp] W(tail_row, :)
  float* hs_part;       // [grid][Kp]
  float* gt_part;       // [grid][Kp]
  double* cost_part;    // [grid]
  int probe;            // print the MMA issuer's wait/issue clocks of CTA 0 (diagnostics)
  int nu;               // units of the streaming ring (2..NU_MAX)
};

struct WPhase2Args {
  int F, Kp, nkb;
  int nchunk, ngroups;  // 128-bin chunks; frame groups (grid = nchunk * ngroups)
  int nblocks;          // ceil(T / 128)
  int ldv;
  long long T;
  const float* V;       // [T][ldv]
  float* Gpart;         // [ngroups][nchunk*128][Kp]
  int nu;
  int probe;            // print the MMA issuer's wait/issue clocks of CTA 0 (diagnostics)
};

__host__ __device__ constexpr size_t phase2_smem_bytes(int nkb, int nu) {
  return (size_t)nkb * 16384 + (size_t)nu * UNIT + 2 * BM * 4 + 64 + 32 * 8 + 1024;
}
// largest ring that fits beside the resident tile
__host__ __device__ constexpr int phase2_units(int nkb, size_t max_smem) {
  int nu = NU_MAX;
  while (nu > 2 && phase2_smem_bytes(nkb, nu) > max_smem) --nu;
  return nu;
}

// Streaming ring, producer and consumer side: unit `pos % nu`, barrier parities follow from the position.
struct RingProducer {
  uint8_t* base; uint64_t* full; uint64_t* empty; uint32_t nu, p;
  __device__ __forceinline__ uint32_t acquire() {
    const uint32_t idx = p % nu;
    umma::mbar_wait(empty + idx, ((p / nu) & 1) ^ 1);
    return idx;
  }
  __device__ __forceinline__ void load_tile(const CUtensorMap* map, int x, int y) {   // box {32, 128}
    const uint32_t idx = acquire();
    umma::mbar_expect_tx(full + idx, UNIT);
    umma::tma_load_2d(base + (size_t)idx * UNIT, map, full + idx, x, y);
    ++p;
  }
  // one 3-D box {32 floats, 16 rows, nkb column blocks}: the slice arrives as nkb consecutive [16 x 128 B] blocks with
  // ONE TMA instruction (eight 2 KB boxes per slice made the second product TMA-issue-bound: ~100 clk per box)
  __device__ __forceinline__ void load_slice(const CUtensorMap* map, int nkb, int y) {
    const uint32_t idx = acquire();
    umma::mbar_expect_tx(full + idx, nkb * SL * 128);
    umma::tma_load_3d(base + (size_t)idx * UNIT, map, full + idx, 0, y, 0);
    ++p;
  }
};
struct RingConsumer {
  uint32_t base_addr; uint64_t* full; uint64_t* empty; uint32_t nu, p;
  __device__ __forceinline__ uint32_t wait() {   // index of the next unit (its data has landed)
    const uint32_t idx = p % nu;
    umma::mbar_wait(full + idx, (p / nu) & 1);
    umma::tc_fence_after();
    return idx;
  }
  __device__ __forceinline__ void release() {    // free again when the MMAs issued so far have completed
    umma::mma_commit(empty + (p % nu));
    ++p;
  }
};

// ------------------------------------------------------------------------------------------------ H phase
__global__ void __launch_bounds__(THREADS, 1)
hphase2_kernel(const __grid_constant__ CUtensorMap mapH, const __grid_constant__ CUtensorMap mapHout,
               const __grid_constant__ CUtensorMap mapWk, const __grid_constant__ CUtensorMap mapWm, const HPhase2Args a) {
  using namespace umma;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  const int nkb = a.nkb, Kp = a.Kp, nh = a.nh;
  const int nu = a.nu;
  uint8_t* Hs = smem;                                     // nkb x [128 x 128 B]   SW128, A of product 1
  uint8_t* Rg = Hs + nkb * 16384;                         // nu x 16 KB streaming ring
  float* vtail_s = (float*)(Rg + (size_t)nu * UNIT);      // [128]
  float* dotp = vtail_s + BM;                             // [128]
  double* red = (double*)(dotp + BM);                     // [8]
  uint64_t* bars = (uint64_t*)(red + 8);
  uint64_t* h_full = bars + 0;
  uint64_t* h_empty = bars + 1;
  uint64_t* num_full = bars + 2;
  uint64_t* num_empty = bars + 3;
  uint64_t* lam_full = bars + 4;                 // [2]
  uint64_t* r_full = bars + 6;                   // [2]
  uint64_t* u_full = bars + 8;                   // [NU_MAX]
  uint64_t* u_empty = bars + 8 + NU_MAX;         // [NU_MAX]
  uint32_t* tmem_slot = (uint32_t*)(bars + 8 + 2 * NU_MAX);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(h_full, 1);
    mbar_init(h_empty, 1);
    mbar_init(num_full, 1);
    mbar_init(num_empty, 2 * BM);
    for (int i = 0; i < 2; ++i) {
      mbar_init(lam_full + i, 1);
      mbar_init(r_full + i, 2 * BM);
    }
    for (int i = 0; i < NU_MAX; ++i) {
      mbar_init(u_full + i, 1);
      mbar_init(u_empty + i, 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const bool upd = a.update != 0;
  auto half_n = [&](int c) { return c == nh - 1 ? a.nlast : HB; };

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0 && (int)blockIdx.x < a.ntiles) {
      tma_prefetch_desc(&mapH);
      tma_prefetch_desc(&mapWk);
      tma_prefetch_desc(&mapWm);
      RingProducer ring{Rg, u_full, u_empty, (uint32_t)nu, 0u};
      int it = 0;
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++it) {
        const int t0 = tile * BM;
        // the dictionary does not depend on the frame tile: the first K-major tiles of this tile go out before the
        // wait for the previous tile's write-back
        const int early = nkb < (nu / 2) ? nkb : (nu / 2);
        for (int ks = 0; ks < early; ++ks) ring.load_tile(&mapWk, ks * KB, 0);
        mbar_wait(h_empty, (it & 1) ^ 1);
        mbar_expect_tx(h_full, nkb * 16384);
        for (int kb = 0; kb < nkb; ++kb) tma_load_2d(Hs + kb * 16384, &mapH, h_full, kb * KB, t0);
        for (int c = 0; c <= nh; ++c) {   // same order as the MMA issuer consumes
          if (c < nh)
            for (int ks = (c == 0 ? early : 0); ks < nkb; ++ks) ring.load_tile(&mapWk, ks * KB, c * HB);
          if (c >= 1 && upd)
            for (int js = 0; js < half_n(c - 1) / SL; ++js) ring.load_slice(&mapWm, nkb, (c - 1) * HB + js * SL);
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (one thread)
    if (lane == 0 && (int)blockIdx.x < a.ntiles) {
      const uint32_t id2 = idesc_tf32(BM, Kp, 0, 1);
      // The issuing thread is on the critical path (one thread, dependent integer chains): descriptors are formed once
      // and then only advanced.  The low 14 bits of a descriptor hold (address >> 4), all addresses are < 256 KB.
      const uint64_t da0 = smem_desc(smem_u32(Hs), 16, 1024);
      const uint64_t dk0 = smem_desc(smem_u32(Rg), 16, 1024);                           // K-major tile in unit 0
      const uint64_t ds0 = smem_desc(smem_u32(Rg), SL * 128, 512, LAYOUT_SW128_32B);    // MN-major slice in unit 0
      RingConsumer ring{smem_u32(Rg), u_full, u_empty, (uint32_t)nu, 0u};
      uint32_t n = 0;
      int it = 0;
      long long p_h = 0, p_a = 0, p_i1 = 0, p_r = 0, p_i2 = 0, p_sw = 0, p_t0 = clock64();
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++it) {
        long long q0 = clock64();
        mbar_wait(h_full, it & 1);
        p_h += clock64() - q0;
        tc_fence_after();
        const uint32_t nbase = n;
        for (int c = 0; c <= nh; ++c) {
          if (c < nh) {  // Lambda(c) = H_tile * W_half(c)'
            const uint32_t m = nbase + c, b = m & 1;
            const uint32_t id1 = idesc_tf32(BM, half_n(c), 0, 0);
            for (int ks = 0; ks < nkb; ++ks) {
              q0 = clock64();
              const uint32_t ui = ring.wait();
              p_a += clock64() - q0;
              q0 = clock64();
              const uint64_t da = da0 + (uint64_t)(ks * (16384 >> 4));
              const uint64_t db = dk0 + (uint64_t)(ui * (UNIT >> 4));
              const uint32_t dt = tmem + LAM_COL + HB * b;
              mma_ss(dt, da, db, id1, ks > 0);
              mma_ss(dt, da + 2, db + 2, id1, 1);
              mma_ss(dt, da + 4, db + 4, id1, 1);
              mma_ss(dt, da + 6, db + 6, id1, 1);
              ring.release();
              p_i1 += clock64() - q0;
            }
            mma_commit(lam_full + b);
          }
          if (c >= 1) {  // Num += R(c-1) * W_half(c-1)
            const uint32_t m = nbase + c - 1, b = m & 1;
            q0 = clock64();
            mbar_wait(r_full + b, (m >> 1) & 1);
            p_r += clock64() - q0;
            tc_fence_after();
            if (upd) {
              if (c == 1) {
                mbar_wait(num_empty, (it & 1) ^ 1);
                tc_fence_after();
              }
              q0 = clock64();
              for (int js = 0; js < half_n(c - 1) / SL; ++js) {
                const long long q1 = clock64();
                const uint32_t ui = ring.wait();
                p_sw += clock64() - q1;
                const uint64_t db = ds0 + (uint64_t)(ui * (UNIT >> 4));
                const uint32_t at = tmem + LAM_COL + HB * b + SL * js;
                mma_ts(tmem, at, db, id2, (c > 1) || (js > 0));
                mma_ts(tmem, at + 8, db + (1024 >> 4), id2, 1);
                ring.release();
              }
              if (c == nh) mma_commit(num_full);
              p_i2 += clock64() - q0;
            }
          }
        }
        n = nbase + nh;
      }
      if (a.probe && blockIdx.x == 0)
        printf("hphase2 probe (MMA issuer, CTA 0; ring %d units, grid %d): total %lld clk, %d tiles x %d halves | "
               "wait h_full %lld, wait K tiles %lld, issue MMA1 %lld, wait r_full %lld, MMA2 (issue + slice waits) %lld of which slice "
               "waits %lld\n",
               nu, (int)gridDim.x, clock64() - p_t0, it, nh, p_h, p_a, p_i1, p_r, p_i2, p_sw);
    }
  } else {
    // ===================================================================== epilogue: two groups of 128 threads
    const int e = (warp - 2) >> 2;           // group: columns [64 e, 64 e + 64) of every Lambda buffer
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
      // the epilogue is issue-bound (two warps per scheduler): whole in-range chunks take a path without per-element
      // predicates or address arithmetic
      auto load_v = [&](int f0, float (&dst)[32]) {
        if (row_ok && f0 + 32 <= a.Fm) {
          const float* pv = vcol + (size_t)f0 * a.ldt;
#pragma unroll
          for (int j = 0; j < 32; ++j) dst[j] = __ldcs(pv + (size_t)j * a.ldt);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) dst[j] = (row_ok && f0 + j < a.Fm) ? __ldg(vcol + (size_t)(f0 + j) * a.ldt) : 0.f;
        }
      };
      float cost_lg = 0.f, cost_lin = 0.f;   // sum v*log2(v/lambda), sum (lambda - v)
      // this thread's chunks of the tile, in order: (half c, 32-column chunk cc) -> first bin 128 c + 64 e + 32 cc
      float vn[32];
      load_v(64 * e, vn);                    // in flight while the first product of the tile runs
      for (int c = 0; c < nh; ++c, ++n) {
        const int Nh = half_n(c);
        const uint32_t b = n & 1;
        mbar_wait(lam_full + b, (n >> 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          const int c0 = 64 * e + 32 * cc;   // column inside the buffer
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = vn[j];
          {
            const int nf = cc == 0 ? HB * c + c0 + 32 : HB * (c + 1) + 64 * e;   // first bin of the next chunk
            if (nf < a.Fm) load_v(nf, vn);
          }
          if (c0 < Nh) {
            uint32_t lam[32];
            tmem_ld32(lane_addr + LAM_COL + HB * b + c0, lam);
            tmem_wait_ld();
            const int f0 = HB * c + c0;
            if (row_ok && f0 + 32 <= a.Fm) {
              if (a.want_cost) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  const float vv = fmaxf(v[j], FLRF);                     // sparse_nmf.m:169
                  const float ll = fmaxf(__uint_as_float(lam[j]), FLRF);  // :167,208
                  const float r = __fdividef(vv, ll);
                  lam[j] = to_tf32_rn(r);
                  cost_lg = fmaf(vv, __log2f(r), cost_lg);                // :250, v*log(v/lambda) = ln2 * v*log2(r)
                  cost_lin += ll - vv;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  const float vv = fmaxf(v[j], FLRF);
                  const float ll = fmaxf(__uint_as_float(lam[j]), FLRF);
                  lam[j] = to_tf32_rn(__fdividef(vv, ll));
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const bool ok = row_ok && (f0 + j < a.Fm);
                const float vv = fmaxf(v[j], FLRF);
                const float ll = fmaxf(__uint_as_float(lam[j]), FLRF);
                const float r = __fdividef(vv, ll);
                lam[j] = ok ? to_tf32_rn(r) : 0u;
                if (a.want_cost && ok) cost_tile += vv * __logf(r) - vv + ll;
              }
            }
            if (upd) tmem_st32(lane_addr + LAM_COL + HB * b + c0, lam);
          }
        }
        if (upd) tmem_wait_st();
        tc_fence_before();
        mbar_arrive(r_full + b);
      }
      // ---- the bin that stays off the tensor cores: Lambda(tail, frame) = W(tail,:) * h (state before the update)
      mbar_wait(h_full, it & 1);
      float r_t = 0.f;
      if (a.tail_row >= 0) {
        float d0 = 0.f;
        for (int kb = e; kb < nkb; kb += 2) {
          const uint8_t* hrow = Hs + kb * 16384 + row * 128;
#pragma unroll
          for (int g4 = 0; g4 < 8; ++g4) {
            const float4 hv = *(const float4*)(hrow + (((g4 ^ row) & 7) << 4));
            const float4 wt = __ldg((const float4*)(a.wtail + kb * KB + 4 * g4));
            d0 += wt.x * h_unbias(hv.x) + wt.y * h_unbias(hv.y) + wt.z * h_unbias(hv.z) + wt.w * h_unbias(hv.w);
          }
        }
        (e == 0 ? vtail_s : dotp)[row] = d0;
        named_bar_sync(1, 2 * BM);
        const float lam_t = fmaxf(vtail_s[row] + dotp[row], FLRF);
        const float vv = fmaxf(row_ok ? __ldg(a.Vt + (size_t)a.tail_row * a.ldt + t0 + row) : 0.f, FLRF);
        r_t = row_ok ? __fdividef(vv, lam_t) : 0.f;
        if (a.want_cost && row_ok && e == 0) cost_tile += vv * __logf(r_t) - vv + lam_t;
        named_bar_sync(1, 2 * BM);
        if (e == 0) vtail_s[row] = vv;
        r_t = __uint_as_float(to_tf32_rn(r_t));
      }
      cost_acc += (double)cost_tile + (double)(0.69314718f * cost_lg + cost_lin);
      if (upd) {
        // ---- H' = H .* Num ./ dph  (sparse_nmf.m:192-195), in place in the shared-memory tile
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
#pragma unroll
          for (int g4 = 0; g4 < 8; ++g4) {
            const int k = kb * KB + 4 * g4;
            const float4 id = __ldg((const float4*)(a.invden + k));
            const float4 wt = __ldg((const float4*)(a.wtail + k));
            float4 x = hv[g4];
            x.x = x.x * (__uint_as_float(num[4 * g4 + 0]) + r_t * wt.x) * id.x;
            x.y = x.y * (__uint_as_float(num[4 * g4 + 1]) + r_t * wt.y) * id.y;
            x.z = x.z * (__uint_as_float(num[4 * g4 + 2]) + r_t * wt.z) * id.z;
            x.w = x.w * (__uint_as_float(num[4 * g4 + 3]) + r_t * wt.w) * id.w;
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
__global__ void __launch_bounds__(THREADS, 1)
wphase2_kernel(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapHk,
               const __grid_constant__ CUtensorMap mapHm, const WPhase2Args a) {
  using namespace umma;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  const int nkb = a.nkb, Kp = a.Kp;
  const int nu = a.nu;
  uint8_t* Wc = smem;                                  // nkb x [128 x 128 B]   resident dictionary rows
  uint8_t* Rg = Wc + nkb * 16384;                      // nu x 16 KB streaming ring (tiles / slices of H')
  uint64_t* bars = (uint64_t*)(Rg + (size_t)nu * UNIT);
  uint64_t* wc_full = bars + 0;
  uint64_t* g_full = bars + 1;
  uint64_t* lam_full = bars + 2;                 // [2]
  uint64_t* r_full = bars + 4;                   // [2]
  uint64_t* u_full = bars + 6;
  uint64_t* u_empty = bars + 6 + NU_MAX;
  uint32_t* tmem_slot = (uint32_t*)(bars + 6 + 2 * NU_MAX);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(wc_full, 1);
    mbar_init(g_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(lam_full + i, 1);
      mbar_init(r_full + i, 2 * BM);
    }
    for (int i = 0; i < NU_MAX; ++i) {
      mbar_init(u_full + i, 1);
      mbar_init(u_empty + i, 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int chunk = blockIdx.x % a.nchunk, grp = blockIdx.x / a.nchunk;
  const int n_my = (a.nblocks > grp) ? (a.nblocks - grp + a.ngroups - 1) / a.ngroups : 0;
  // N of block i of this CTA (frames beyond T read as zero rows; N stays a multiple of 16)
  auto block_n = [&](int i) {
    const long long t0 = (long long)(grp + i * a.ngroups) * HB;
    const long long left = a.T - t0;
    return left >= HB ? HB : (int)((left + 15) / 16 * 16);
  };

  if (warp == 0) {
    if (lane == 0 && n_my > 0) {
      tma_prefetch_desc(&mapW);
      tma_prefetch_desc(&mapHk);
      tma_prefetch_desc(&mapHm);
      mbar_expect_tx(wc_full, nkb * 16384);
      for (int kb = 0; kb < nkb; ++kb) tma_load_2d(Wc + kb * 16384, &mapW, wc_full, kb * KB, chunk * BM);
      RingProducer ring{Rg, u_full, u_empty, (uint32_t)nu, 0u};
      for (int i = 0; i <= n_my; ++i) {   // same order as the MMA issuer consumes
        if (i < n_my)
          for (int ks = 0; ks < nkb; ++ks) ring.load_tile(&mapHk, ks * KB, (grp + i * a.ngroups) * HB);
        if (i >= 1)
          for (int js = 0; js < block_n(i - 1) / SL; ++js)
            ring.load_slice(&mapHm, nkb, (grp + (i - 1) * a.ngroups) * HB + js * SL);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && n_my > 0) {
      const uint32_t id4 = idesc_tf32(BM, Kp, 0, 1);
      const uint64_t da0 = smem_desc(smem_u32(Wc), 16, 1024);
      const uint64_t dk0 = smem_desc(smem_u32(Rg), 16, 1024);
      const uint64_t ds0 = smem_desc(smem_u32(Rg), SL * 128, 512, LAYOUT_SW128_32B);
      RingConsumer ring{smem_u32(Rg), u_full, u_empty, (uint32_t)nu, 0u};
      mbar_wait(wc_full, 0);
      tc_fence_after();
      long long p_a = 0, p_i1 = 0, p_r = 0, p_i2 = 0, p_sw = 0, q0, p_t0 = clock64();
      for (int i = 0; i <= n_my; ++i) {
        if (i < n_my) {  // Lambda(i) = W_rows * H'_block(i)'
          const uint32_t b = i & 1;
          const uint32_t id3 = idesc_tf32(BM, block_n(i), 0, 0);
          for (int ks = 0; ks < nkb; ++ks) {
            q0 = clock64();
            const uint32_t ui = ring.wait();
            p_a += clock64() - q0;
            q0 = clock64();
            const uint64_t da = da0 + (uint64_t)(ks * (16384 >> 4));
            const uint64_t db = dk0 + (uint64_t)(ui * (UNIT >> 4));
            const uint32_t dt = tmem + LAM_COL + HB * b;
            mma_ss(dt, da, db, id3, ks > 0);
            mma_ss(dt, da + 2, db + 2, id3, 1);
            mma_ss(dt, da + 4, db + 4, id3, 1);
            mma_ss(dt, da + 6, db + 6, id3, 1);
            ring.release();
            p_i1 += clock64() - q0;
          }
          mma_commit(lam_full + b);
        }
        if (i >= 1) {  // G += R(i-1) * H'_block(i-1)
          const int m = i - 1;
          const uint32_t b = m & 1;
          q0 = clock64();
          mbar_wait(r_full + b, (m >> 1) & 1);
          p_r += clock64() - q0;
          tc_fence_after();
          q0 = clock64();
          for (int js = 0; js < block_n(m) / SL; ++js) {
            const long long q1 = clock64();
            const uint32_t ui = ring.wait();
            p_sw += clock64() - q1;
            const uint64_t db = ds0 + (uint64_t)(ui * (UNIT >> 4));
            const uint32_t at = tmem + LAM_COL + HB * b + SL * js;
            mma_ts(tmem, at, db, id4, (m > 0) || (js > 0));
            mma_ts(tmem, at + 8, db + (1024 >> 4), id4, 1);
            ring.release();
          }
          p_i2 += clock64() - q0;
          if (m == n_my - 1) mma_commit(g_full);
        }
      }
      if (a.probe && blockIdx.x == 0)
        printf("wphase2 probe (MMA issuer, CTA 0; ring %d units, grid %d): total %lld clk, %d blocks | wait K tiles %lld, "
               "issue MMA1 %lld, wait r_full %lld, MMA2 (issue + slice waits) %lld of which slice waits %lld\n",
               nu, (int)gridDim.x, clock64() - p_t0, n_my, p_a, p_i1, p_r, p_i2, p_sw);
    }
  } else {
    const int e = (warp - 2) >> 2, q = warp & 3;
    const int row = 32 * q + lane;
    const int f = chunk * BM + row;
    const bool f_ok = f < a.F;
    const uint32_t lane_addr = tmem + ((uint32_t)(32 * q) << 16);
    // this thread's 32-frame chunks, in order: (block i, chunk cc) -> first frame (grp + i ngroups) 128 + 64 e + 32 cc
    auto load_v = [&](long long tf, float (&dst)[32]) {
      if (f_ok && tf + 32 <= a.T) {
        const float* pv = a.V + (size_t)tf * a.ldv + f;
#pragma unroll
        for (int j = 0; j < 32; ++j) dst[j] = __ldcs(pv + (size_t)j * a.ldv);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) dst[j] = (f_ok && tf + j < a.T) ? __ldg(a.V + (size_t)(tf + j) * a.ldv + f) : 0.f;
      }
    };
    float vn[32];
    if (n_my > 0) load_v((long long)grp * HB + 64 * e, vn);
    for (int i = 0; i < n_my; ++i) {
      const int Nb = block_n(i);
      const long long tb = (long long)(grp + i * a.ngroups) * HB;
      const uint32_t b = i & 1;
      mbar_wait(lam_full + b, (i >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int c0 = 64 * e + 32 * cc;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = vn[j];
        {
          const long long nf = cc == 0 ? tb + c0 + 32 : (long long)(grp + (i + 1) * a.ngroups) * HB + 64 * e;
          if ((cc == 0 || i + 1 < n_my) && nf < a.T) load_v(nf, vn);
        }
        if (c0 < Nb) {
          uint32_t lam[32];
          tmem_ld32(lane_addr + LAM_COL + HB * b + c0, lam);
          tmem_wait_ld();
          if (f_ok && tb + c0 + 32 <= a.T) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              lam[j] = to_tf32_rn(__fdividef(fmaxf(v[j], FLRF), fmaxf(__uint_as_float(lam[j]), FLRF)));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const bool ok = f_ok && (tb + c0 + j < a.T);
              const float r = __fdividef(fmaxf(v[j], FLRF), fmaxf(__uint_as_float(lam[j]), FLRF));
              lam[j] = ok ? to_tf32_rn(r) : 0u;
            }
          }
          tmem_st32(lane_addr + LAM_COL + HB * b + c0, lam);
        }
      }
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(r_full + b);
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
