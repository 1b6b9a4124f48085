// mma_probe.cu -- how many clocks does one tcgen05.mma take on this GPU, per shape / operand source / kind?
// (tool, not product: answers "what bounds tc_conv_kernel", see profiles/r2_mma_probe.md)
//
// One CTA per SM; one thread issues `iters` x 4 K-steps x `terms` MMAs on static operands (no TMA, no converter
// warps, no epilogue), commits to an mbarrier and waits; clock64 around it.  Variants:
//   kind: tf32 (K = 8 per MMA) / bf16 (K = 16)      A: tensor memory / shared memory      N: 64 / 128 / 256
//   same: every MMA accumulates into the same D / alt: two D buffers alternate per K block
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/mma_probe tools/probes/mma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned s32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long desc_k128(unsigned saddr) {
  return (unsigned long long)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mbar_init(unsigned bar, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  unsigned done;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}

template <int KIND, bool ATM>   // KIND 0 = tf32, 1 = bf16 (kind::f16)
__device__ __forceinline__ void mma(unsigned d, unsigned a_tmem, unsigned long long adesc, unsigned long long bdesc,
                                    unsigned idesc, unsigned acc) {
  if (KIND == 0) {
    if (ATM) asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
    else asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
  } else {
    if (ATM) asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
    else asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
  }
}

template <int KIND, bool ATM, int N>
__global__ void __launch_bounds__(128, 1) probe(int iters, int terms, int alt, long long *cycles) {
  extern __shared__ unsigned char raw[];
  const unsigned base = (s32(raw) + 1023u) & ~1023u;
  unsigned char *g = raw + (base - s32(raw));
  // [A: 2 x 16 KB][B: 2 x N*128 B][barrier][tmem slot]
  const unsigned a_s = base, b_s = base + 32768, bar = b_s + 2 * N * 128, slot = bar + 8;
  for (int i = threadIdx.x; i < (32768 + 2 * N * 128) / 4; i += blockDim.x)
    reinterpret_cast<unsigned *>(g)[i] = KIND == 0 ? 0x3F800000u + ((i * 2654435761u) & 0x007FE000u)      // ~1.x fp32 / tf32
                                                   : 0x3F803F80u ^ ((i * 2654435761u) & 0x007F007Fu);     // pairs of ~1.x bf16
  if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = *reinterpret_cast<volatile unsigned *>(g + (slot - base));
  // A operand in tensor memory: columns [2N, 2N + 64) (hi | lo), every lane, some finite pattern
  if (ATM) {
    const unsigned ta = tmem + ((unsigned)((threadIdx.x >> 5) * 32) << 16) + (unsigned)(2 * N > 448 ? 448 : 2 * N);
    for (int c = 0; c < 64; ++c) {
      const unsigned v = KIND == 0 ? 0x3F800000u + ((c * 40503u + threadIdx.x * 977u) & 0x007FE000u) : 0x3F803F80u;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(ta + c), "r"(v) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x == 0) {
    const unsigned fmt = KIND == 0 ? 2u : 1u;      // tf32 / bf16
    const unsigned idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((unsigned)(N >> 3) << 17) | ((128u >> 4) << 24);
    const unsigned acol = (unsigned)(2 * N > 448 ? 448 : 2 * N);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      // N = 256 has room for one accumulator only (512 columns): alt is ignored there
      const unsigned d = tmem + ((alt && N <= 128) ? (unsigned)((it & 1) * N) : 0u);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        for (int t = 0; t < terms; ++t) {
          const unsigned long long bd = desc_k128(b_s + (t == 1 ? N * 128 : 0)) + (unsigned long long)(2 * k);
          const unsigned long long ad = desc_k128(a_s + (t == 0 ? 16384 : 0)) + (unsigned long long)(2 * k);
          const unsigned at = tmem + acol + (t == 0 ? 32u : 0u) + 8u * k;   // 32 bytes of A per row and K step for both kinds
          mma<KIND, ATM>(d, at, ad, bd, idesc, (it | k | t) != 0);
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int KIND, bool ATM, int N>
static void run(const char *name, int ctas, int iters, int terms, int alt) {
  const int smem = 32768 + 2 * N * 128 + 64 + 1024;
  cudaFuncSetAttribute(probe<KIND, ATM, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long *dc;
  cudaMalloc(&dc, sizeof(long long) * ctas);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<KIND, ATM, N><<<ctas, 128, smem>>>(iters / 8 + 1, terms, alt, dc);      // warm-up
  cudaEventRecord(e0);
  probe<KIND, ATM, N><<<ctas, 128, smem>>>(iters, terms, alt, dc);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) { printf("{\"case\": \"%s\", \"error\": \"%s\"}\n", name, cudaGetErrorString(err)); exit(1); }
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  long long *hc = (long long *)malloc(sizeof(long long) * ctas);
  cudaMemcpy(hc, dc, sizeof(long long) * ctas, cudaMemcpyDeviceToHost);
  double sum = 0;
  for (int i = 0; i < ctas; ++i) sum += (double)hc[i];
  const double n_mma = (double)iters * 4 * terms;
  const double kper = KIND == 0 ? 8 : 16;
  const double flops = n_mma * 2.0 * 128 * N * kper * ctas;
  printf("{\"case\": \"%s\", \"kind\": \"%s\", \"a\": \"%s\", \"N\": %d, \"terms\": %d, \"alt\": %d, \"ctas\": %d, "
         "\"clk_per_mma\": %.1f, \"ms\": %.3f, \"tflops\": %.1f}\n",
         name, KIND == 0 ? "tf32" : "bf16", ATM ? "tmem" : "smem", N, terms, alt, ctas, sum / ctas / n_mma, ms,
         flops / (ms * 1e-3) / 1e12);
  fflush(stdout);
  cudaFree(dc); free(hc);
}

int main(int argc, char **argv) {
  int sms = 148;
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, 0) == cudaSuccess) sms = p.multiProcessorCount;
  const int it = 4096;
  for (int ctas : {1, sms}) {
    run<0, true, 64>("tf32 A=tmem N=64", ctas, it, 3, 1);
    run<0, true, 128>("tf32 A=tmem N=128", ctas, it, 3, 1);
    run<0, true, 128>("tf32 A=tmem N=128 same D", ctas, it, 3, 0);
    run<0, true, 128>("tf32 A=tmem N=128 1 term", ctas, it, 1, 1);
    run<0, true, 256>("tf32 A=tmem N=256", ctas, it, 3, 0);
    run<0, false, 64>("tf32 A=smem N=64", ctas, it, 3, 1);
    run<0, false, 128>("tf32 A=smem N=128", ctas, it, 3, 1);
    run<0, false, 256>("tf32 A=smem N=256", ctas, it, 3, 0);
    run<1, true, 128>("bf16 A=tmem N=128", ctas, it, 3, 1);
    run<1, true, 256>("bf16 A=tmem N=256", ctas, it, 3, 0);
    run<1, false, 128>("bf16 A=smem N=128", ctas, it, 3, 1);
    run<1, false, 256>("bf16 A=smem N=256", ctas, it, 3, 0);
  }
  return 0;
}
