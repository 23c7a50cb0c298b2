// Device helpers shared by the per-hop MU kernels (online_fast.cu, online_ms.cu): Newton reciprocal, table-driven log,
// mbarrier + st.async pushes over distributed shared memory.
#pragma once
#include "online.cuh"

namespace snmfnat {

__device__ __forceinline__ double fast_rcp(double x) {  // x > 0, normal.  <= 1 ulp after two Newton steps
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  return y;
}

// log(x) for positive normal x: x = 2^e * m, m in [1,2); c = centre of m's 1/128 bucket; z = m/c - 1, |z| <= 2^-8;
// log x = e ln2 + log c + log1p(z) with a degree-7 series.  tab[i] = {1/c_i, log c_i}.
__device__ __forceinline__ double fast_log(double x, const double2* __restrict__ tab) {
  const int hi = __double2hiint(x);
  const int e = (hi >> 20) - 1023;
  const int idx = (hi >> 13) & 127;
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
  const double2 t = tab[idx];
  const double z = fma(m, t.x, -1.0);
  double p = fma(z, 1.0 / 7.0, -1.0 / 6.0);
  p = fma(p, z, 0.2);
  p = fma(p, z, -0.25);
  p = fma(p, z, 1.0 / 3.0);
  p = fma(p, z, -0.5);
  const double l = fma(p, z * z, z);
  return fma((double)e, 0.693147180559945309417232, t.y + l);
}

// mbarrier + st.async (push over distributed shared memory, completion counted in bytes on the receiver's barrier)
__device__ __forceinline__ unsigned hf_mapa(unsigned addr, unsigned rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void hf_st_async(unsigned raddr, double v, unsigned rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
               :: "r"(raddr), "l"(__double_as_longlong(v)), "r"(rbar) : "memory");
}
__device__ __forceinline__ void hf_mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void hf_mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void hf_mbar_wait(unsigned bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}


// Bounded mbarrier wait: a protocol bug must surface as a trapped kernel (an error at the next sync), never as a hung GPU.
__device__ __forceinline__ void hf_mbar_wait_bounded(unsigned bar, unsigned parity) {
  unsigned ok;
  long long t0 = 0;
  for (unsigned spin = 0;; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
    if ((spin & 1023u) == 1023u) {
      const long long t = clock64();
      if (t0 == 0) t0 = t;
      else if (t - t0 > 6000000000LL) {
        printf("snmfnat: mbarrier timeout (block %d thread %d bar %u parity %u)\n", (int)blockIdx.x, (int)threadIdx.x, bar, parity);
        __trap();
      }
    }
  }
}

// One bulk copy of `bytes` (multiple of 16, 16-byte aligned) from this CTA's shared memory into the shared memory of a CTA
// of the cluster; the bytes are counted on the RECEIVER's mbarrier (cp.async.bulk, async proxy -> SASS UBLKCP).
__device__ __forceinline__ void hf_bulk_push(unsigned rdst, unsigned src, unsigned bytes, unsigned rbar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(rdst), "r"(src), "r"(bytes), "r"(rbar) : "memory");
}
// generic-proxy writes to shared memory become visible to the async proxy (call before a barrier that precedes a bulk copy)
__device__ __forceinline__ void hf_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// D(8x8) += A(8x4) * B(4x8) on the FP64 tensor cores.  lane = 4*i + j:  a = A[i][j],  b = B[j][i],  c0/c1 = C[i][2j], C[i][2j+1]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// the 128-entry {1/c, log c} table of fast_log on the context's device (created on first use)
const double2* log_table(snmfnat_ctx* ctx);

}  // namespace snmfnat
