// tc_rate.cu -- issue-rate probe: cycles per tcgen05.mma kind::tf32 for several shapes (bring-up tool).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" :: "r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred P1;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DN;\n\tbra WL;\n\tDN:\n\t}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// variant: 0 = TS same D, 1 = TS rotating over 4 D regions, 2 = SS same D; n_issuers warps issue concurrently
__global__ void __launch_bounds__(128, 1) rate(int M, int N, int variant, int n_mma, long long* out, int n_issuers) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tbase_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 48 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.001f * (i % 97);
  if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&mbar)), "r"(n_issuers) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tbase_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tbase_s;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  // B: [N][8] K-major: LBO = N*16 (next k core matrix), SBO = 128
  const uint64_t bdesc = make_desc(smem_u32(smem), (uint32_t)N * 16, 128);
  const uint64_t adesc = make_desc(smem_u32(smem + 16384), (uint32_t)M * 16, 128);
  if ((tid & 31) == 0 && warp < n_issuers) {
    const uint32_t dbase = tb + (uint32_t)(warp * 64);
    const uint32_t abase = tb + 256 + (uint32_t)(warp * 32);
    long long t0 = clock64();
    for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t d = dbase + ((variant == 1) ? (uint32_t)((j & 1) * 32) : 0u);
        if (variant == 2) mma_ss(d, adesc, bdesc, idesc, 1u);
        else mma_ts(d, abase + (uint32_t)((j & 3) * 8), bdesc, idesc, 1u);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
    long long t1 = clock64();
    mbar_wait(&mbar, 0);
    long long t2 = clock64();
    if (warp == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tb) : "memory");
}
#include <cstring>
// --peak: dense TF32 tcgen05 throughput of the whole chip (one CTA per SM, one issuer, back-to-back accumulating MMAs),
// timed with CUDA events; prints one JSON line (-> profiles/tf32_peak.json: the denominator of bench.py's tensor_tf32 fraction).
static int peak_mode() {
  long long* d; CK(cudaMalloc(&d, 16));
  CK(cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  struct { int M, N, variant; const char* name; } cases[] = {{128, 256, 0, "M128 N256 A-in-TMEM"}, {128, 256, 2, "M128 N256 A,B in smem"},
                                                              {128, 32, 0, "M128 N32 A-in-TMEM (the chain pass shape)"},
                                                              {64, 64, 2, "M64 N64 A,B in smem (the dW shape)"}};
  double best = 0.0;
  printf("{\"device\": \"%s\", \"sms\": %d, \"cases\": [", prop.name, sms);
  for (int c = 0; c < 4; ++c) {
    const int n_mma = 200000;
    rate<<<sms, 128, 64 * 1024>>>(cases[c].M, cases[c].N, cases[c].variant, 4096, d, 1); CK(cudaDeviceSynchronize());
    float ms_best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
      CK(cudaEventRecord(e0));
      rate<<<sms, 128, 64 * 1024>>>(cases[c].M, cases[c].N, cases[c].variant, n_mma, d, 1);
      CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (ms < ms_best) ms_best = ms;
    }
    const double tflops = 2.0 * cases[c].M * cases[c].N * 8.0 * (double)n_mma * sms / (ms_best * 1e-3) / 1e12;
    if (c < 2 && tflops > best) best = tflops;
    printf("%s{\"shape\": \"%s\", \"ms\": %.4f, \"tflops\": %.1f}", c ? ", " : "", cases[c].name, ms_best, tflops);
  }
  printf("], \"tf32_tflops\": %.1f, \"source\": \"tools/tc_rate --peak on this B200: tcgen05.mma kind::tf32 M=128 N=256 K=8 back to back, one CTA per SM, CUDA events, best of 5\"}\n", best);
  return 0;
}
int main(int argc, char** argv) {
  if (argc > 1 && !strcmp(argv[1], "--peak")) return peak_mode();
  long long* d; CK(cudaMalloc(&d, 16));
  CK(cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  const char* vn[3] = {"TS same-D", "TS 2 D regions", "SS same-D"};
  int shapes[][2] = {{128, 16}, {128, 32}, {128, 64}, {128, 128}, {128, 256}, {64, 32}, {64, 64}};
  for (auto& s : shapes)
    for (int v = 0; v < 3; ++v)
      for (int ni = 1; ni <= 4; ni *= 2) {
        if (v == 1 && s[1] > 32) continue;
        if (ni > 1 && (s[1] > 64 || v == 2)) continue;
        long long h1[2], h2[2];
        rate<<<1, 128, 64 * 1024>>>(s[0], s[1], v, 64, d, ni); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h1, d, 16, cudaMemcpyDeviceToHost));
        rate<<<1, 128, 64 * 1024>>>(s[0], s[1], v, 576, d, ni); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h2, d, 16, cudaMemcpyDeviceToHost));
        const double per = (double)(h2[1] - h1[1]) / 512;
        printf("M=%3d N=%3d K=8 %-15s issuers=%d: issue %.1f, complete %.1f cyc per mma per issuer -> %.1f cyc/mma aggregate (%.0f MAC/cyc)\n", s[0], s[1], vn[v], ni,
               (double)(h2[0] - h1[0]) / 512, per, per / ni, (double)s[0] * s[1] * 8 * ni / per);
      }
  return 0;
}
