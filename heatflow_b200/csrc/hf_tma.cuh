// heatflow_b200 - TMA bulk-copy / mbarrier helpers and the shared-memory stage layout of the streaming
// PCG kernels (hf_pcg.cu: one launch per iteration, hf_stream.cu: one cooperative launch per solve).
#pragma once
#include "hf_ctx.cuh"

__device__ __forceinline__ unsigned hf_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// TMA bulk copy global -> shared (cp.async.bulk, SASS UBLKCP), completion counted on an mbarrier.
// Addresses and size must be multiples of 16 bytes.
__device__ __forceinline__ void hf_bulk_g2s(void* dst_smem, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   hf_smem_u32(dst_smem)),
               "l"(src), "r"(bytes), "r"(hf_smem_u32(bar))
               : "memory");
}
// same with an L2 eviction-priority hint (createpolicy): the operator is streamed once per iteration
// (evict_first) so that the PCG vectors, re-read every iteration, keep their L2 lines
__device__ __forceinline__ void hf_bulk_g2s_hint(void* dst_smem, const void* src, unsigned bytes, unsigned long long* bar,
                                                 unsigned long long policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          hf_smem_u32(dst_smem)),
      "l"(src), "r"(bytes), "r"(hf_smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ unsigned long long hf_policy_evict_first() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ unsigned long long hf_policy_evict_last() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void hf_mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(hf_smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void hf_mbar_expect(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(hf_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void hf_mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "HF_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra HF_DONE;\n"
      "bra HF_WAIT;\n"
      "HF_DONE:\n"
      "}\n" ::"r"(hf_smem_u32(bar)),
      "r"(parity)
      : "memory");
}

#define HF_BULK_PIECE 16384u   // bytes per cp.async.bulk

// Shared-memory stage of one chunk (byte offsets; every block is a multiple of 16 bytes):
//   val[mat_cap] f64 | x[R] | r[R] | q[R] | p[R] | halo p[halo_cap] | lcol[mat_cap] u16 | slice_ptr[R/32 + 4] i32
// After phase 1 the r block holds r_n and the p block + halo block hold p_n (own rows, then halo).
struct IterStage {
  int mat_cap, halo_cap, nstages;
  unsigned stage_bytes;
};

template <int R>
__device__ __forceinline__ void hf_issue_chunk(const PatchView& A, int ch, int e0, int e1, unsigned char* st, int mat_cap,
                                               int halo_cap, const double* x, const double* ro, const double* po,
                                               const double* qo, unsigned long long* bar) {
  constexpr int SPC = R / HF_SLICE;
  const unsigned n = (unsigned)(e1 - e0);
  const unsigned vb = (unsigned)R * 8u;
  unsigned char* sx = st + (size_t)mat_cap * 8;
  unsigned char* scol = sx + 4 * (size_t)vb + (size_t)halo_cap * 8;
  unsigned char* sptr = scol + (size_t)mat_cap * 2;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the stage was last touched by ordinary loads/stores
  hf_mbar_expect(bar, n * 10u + 4u * vb + (SPC + 4) * 4u);
  const size_t lo = (size_t)ch * R;
  const unsigned long long keep = hf_policy_evict_last(), stream = hf_policy_evict_first();
  hf_bulk_g2s_hint(sx, x + lo, vb, bar, keep);
  hf_bulk_g2s_hint(sx + vb, ro + lo, vb, bar, keep);
  hf_bulk_g2s_hint(sx + 2 * vb, qo + lo, vb, bar, keep);
  hf_bulk_g2s_hint(sx + 3 * vb, po + lo, vb, bar, keep);
  hf_bulk_g2s(sptr, A.slice_ptr + (size_t)ch * SPC, (SPC + 4) * 4u, bar);   // slice_ptr is padded to whole chunks + 4
  for (unsigned off = 0; off < n * 8u; off += HF_BULK_PIECE)
    hf_bulk_g2s_hint(st + off, reinterpret_cast<const unsigned char*>(A.val + e0) + off, min(HF_BULK_PIECE, n * 8u - off), bar, stream);
  for (unsigned off = 0; off < n * 2u; off += HF_BULK_PIECE)
    hf_bulk_g2s_hint(scol + off, reinterpret_cast<const unsigned char*>(A.lcol + e0) + off, min(HF_BULK_PIECE, n * 2u - off), bar,
                     stream);
}

