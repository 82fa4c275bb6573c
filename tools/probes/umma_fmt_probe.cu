// Probe (tuning tool, not part of the library): does tcgen05.mma kind::f16 accept a given (A format, B format) pair?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_fmt_probe umma_fmt_probe.cu && ./umma_fmt_probe <a_fmt> <b_fmt> [N]
// Formats: 0 = fp16, 1 = bf16.  A = ones (128 x 16), B = twos (N x 16): every accumulator entry must read 32.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) probe(uint32_t idesc, unsigned short a_one, unsigned short b_two, float* out, int N) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* base = (unsigned char*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  unsigned short* sA = (unsigned short*)base;                 // 128 rows x 64 columns (128 B rows, swizzle irrelevant for constants)
  unsigned short* sB = (unsigned short*)(base + 128 * 128);   // N rows x 64 columns
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bar;
  for (int i = threadIdx.x; i < 128 * 64; i += 128) sA[i] = a_one;
  for (int i = threadIdx.x; i < N * 64; i += 128) sB[i] = b_two;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  auto desc = [](uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
  };
  if (threadIdx.x == 0) {
    const uint64_t da = desc(smem_u32(sA)), db = desc(smem_u32(sB));
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem),
        "l"(da), "l"(db), "r"(idesc), "r"(0u)
        : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // wait for the commit
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n.reg .pred q;\nmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\nselp.u32 %0, 1, 0, q;\n}\n"
                   : "=r"(done)
                   : "r"(smem_u32(&bar)), "r"(0u)
                   : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r[8];
  const uint32_t taddr = tmem + ((uint32_t)(threadIdx.x & ~31) << 16);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  out[threadIdx.x] = __uint_as_float(r[0]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

int main(int argc, char** argv) {
  const uint32_t a_fmt = argc > 1 ? atoi(argv[1]) : 0, b_fmt = argc > 2 ? atoi(argv[2]) : 0;
  const int N = argc > 3 ? atoi(argv[3]) : 64;
  const uint32_t idesc = (1u << 4) | (a_fmt << 7) | (b_fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const unsigned short one16[2] = {0x3C00, 0x3F80}, two16[2] = {0x4000, 0x4000};   // fp16 / bf16 encodings of 1.0 and 2.0
  float* out;
  cudaMalloc(&out, 128 * sizeof(float));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 128 + 256 * 128 + 1024);
  probe<<<1, 128, 128 * 128 + 256 * 128 + 1024>>>(idesc, one16[a_fmt & 1], two16[b_fmt & 1], out, N);
  cudaError_t e = cudaDeviceSynchronize();
  float h[128] = {0};
  if (e == cudaSuccess) cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("a_fmt=%u b_fmt=%u N=%d idesc=0x%08x -> %s, acc[0]=%g acc[127]=%g (expect 32)\n", a_fmt, b_fmt, N, idesc, cudaGetErrorString(e), h[0],
         h[127]);
  return e == cudaSuccess ? 0 : 1;
}
